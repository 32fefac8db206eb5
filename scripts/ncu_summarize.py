#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON: one entry per captured kernel
with the metrics the design discussion uses.  Usage: ncu_summarize.py in.ncu-rep out.json [cells]"""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__icc_request_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cells = float(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")]}
        for i, k in enumerate(hdr):
            if k in KEYS or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if "issue_stalled" in k:
                    e.setdefault("stall_cycles_per_issue", {})[k.split("issue_stalled_")[1].split("_per_issue")[0]] = round(v, 3)
                else:
                    e[k] = [v, units[i]]
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = e["dram__bytes_read.sum"]; wr = e["dram__bytes_write.sum"]
            tot = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
            e["dram_bytes_total"] = tot
            t = e["gpu__time_duration.sum"]
            sec = t[0] * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[t[1]]
            e["dram_GBps_under_ncu"] = tot / sec / 1e9
            if cells:
                e["dram_bytes_per_cell"] = tot / cells
                e["thread_inst_per_cell"] = e["smsp__inst_executed.sum"][0] * 32 / cells
        except Exception as ex:   # noqa
            e["note"] = "derived metrics unavailable: %s" % ex
        res.append(e)
    json.dump({"source": rep, "cells_per_launch": cells, "kernels": res}, open(out, "w"), indent=1)
    for e in res:
        print(e["kernel"][:60], e.get("gpu__time_duration.sum"), "B/cell", round(e.get("dram_bytes_per_cell", 0), 1),
              "inst/cell", round(e.get("thread_inst_per_cell", 0)), "issue%", e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", [0])[0],
              "fp64%", e.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", [0])[0])


if __name__ == "__main__":
    main()
