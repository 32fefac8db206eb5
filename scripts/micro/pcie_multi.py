"""What the HOST side of a multi-GPU box gives the host-layout (BGC_MEM_HOST_FORTRAN) path: pinned
H2D / D2H / both, on all ranks AT ONCE (torchrun, one rank per GPU), with the pinned buffers
(a) wherever the default policy puts them and (b) first-touched by a thread bound to the CPUs of
the GPU's NUMA node.  Prints one JSON line on rank 0 (per-rank and aggregate GB/s) plus the
topology the numbers belong to.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/micro/pcie_multi.py
"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist


def numa_of_gpu(i):
    try:
        bdf = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    except Exception:
        bdf = None
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(i), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
        bdf = out.lower().replace("00000000:", "0000:")
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read()), bdf
    except Exception as e:   # noqa: BLE001
        return -1, str(bdf or e)


def cpus_of_node(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return []


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 27   # doubles: 1 GiB per buffer
    node, bdf = numa_of_gpu(local)
    res = {"rank": rank, "gpu_numa_node": node, "bdf": bdf, "affinity": len(os.sched_getaffinity(0))}
    d1 = torch.empty(n, dtype=torch.float64, device="cuda")
    d2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for mode in ("default", "bound"):
        if mode == "bound":
            cpus = cpus_of_node(node) if node >= 0 else []
            if not cpus:
                res["bound"] = "no NUMA information"
                break
            try:
                os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
            except Exception as e:   # noqa: BLE001
                res["bound"] = "sched_setaffinity failed: %s" % e
                break
        h1 = torch.empty(n, dtype=torch.float64, pin_memory=True); h1.zero_()
        h2 = torch.empty(n, dtype=torch.float64, pin_memory=True); h2.zero_()

        def run(up, down, reps=4):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                if up:
                    with torch.cuda.stream(s1):
                        d1.copy_(h1, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        h2.copy_(d2, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            return reps * n * 8 / dt / 1e9
        run(True, True, 1)
        res[mode] = {"h2d": run(True, False), "d2h": run(False, True), "both_each": run(True, True)}
        # host memset / memcpy bandwidth of this rank while all ranks do the same (what a host-side
        # unpack of compacted diagnostics would have to live with)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            h1.zero_()
        res[mode]["host_memset"] = 2 * n * 8 / (time.perf_counter() - t0) / 1e9
        del h1, h2
    allres = [None] * world
    if world > 1:
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        out = {"world": world, "ranks": allres}
        for mode in ("default", "bound"):
            rs = [r[mode] for r in allres if isinstance(r.get(mode), dict)]
            if rs:
                out["aggregate_" + mode] = {k: sum(r[k] for r in rs) * (2 if k == "both_each" else 1) for k in rs[0]}
        try:
            out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-3000:]
            out["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
            out["cpus"] = os.cpu_count()
            out["mem_gb"] = int(open("/proc/meminfo").readline().split()[1]) / 1e6
        except Exception:
            pass
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
