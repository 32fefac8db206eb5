#!/usr/bin/env python
"""profiles/traffic.json from a summary made by scripts/ncu_summarize.py.

bench.py cannot run ncu inside its timed region, so `roofline.traffic` and `roofline.ncu_counters`
of the bench line are read from this file: the DRAM bytes per cell and the pipe / issue / stall
counters of ONE `ncu --set full --clock-control none` launch of every kernel on the EC60to30 mesh
(scripts/gpu_round.sh), stamped with the commit the capture was taken at.

Usage: make_traffic_json.py profiles/ncu_r02_vNN_summary.json <commit> [out.json]
"""
import json
import re
import sys


def short_name(kernel):
    """'void unnamed>::eco_columns_kernel<2, 256, 1>(EcoArgs)' -> ('eco_columns_kernel<2, 256, 1>', 'eco_columns_kernel')"""
    m = re.search(r"([A-Za-z0-9_]+)(<[^>]*>)?\(", kernel)
    base = m.group(1)
    return base + (m.group(2) or ""), base


def main():
    src, commit = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else "profiles/traffic.json"
    summ = json.load(open(src))
    per_kernel, counters = {}, {}
    for k in summ["kernels"]:
        full, base = short_name(k["kernel"])
        per_kernel[full] = k["dram_bytes_per_cell"]
        stalls = sorted(k["stall_cycles_per_issue"].items(), key=lambda kv: -kv[1])
        stalls = [(n, v) for n, v in stalls if n != "selected"][:4]
        counters[base] = {
            "fp64_pipe_active_pct": round(k["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0], 1),
            "issue_slots_busy_pct": round(k["smsp__issue_active.avg.pct_of_peak_sustained_active"][0], 1),
            "thread_instructions_per_cell": int(round(k["thread_inst_per_cell"])),
            "dram_bytes_per_cell": round(k["dram_bytes_per_cell"], 1),
            "registers_per_thread": int(k["launch__registers_per_thread"][0]),
            "warps_active_pct": round(k["sm__warps_active.avg.pct_of_peak_sustained_active"][0], 1),
            "top_stalls_cycles_per_issue": dict(stalls),
        }
    eco = [v for k, v in per_kernel.items() if k.startswith("eco_columns_kernel")]
    what = "%s (ncu --set full --clock-control none, one launch each on the EC60to30 mesh)" % src
    doc = {
        "source": "%s (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one launch of "
                  "every kernel on the EC60to30 mesh, scripts/gpu_round.sh)" % src,
        "commit": commit,
        "eco_columns_kernel_bytes_per_cell": eco[0] if eco else None,
        "per_kernel_bytes_per_cell": per_kernel,
        "ncu_counters": counters,
        "ncu_counters_source": what,
    }
    json.dump(doc, open(out, "w"), indent=1)
    print("wrote", out, "eco bytes/cell", doc["eco_columns_kernel_bytes_per_cell"])


if __name__ == "__main__":
    main()
