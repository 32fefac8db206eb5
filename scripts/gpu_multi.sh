#!/bin/bash
# One visit to an N-GPU box: the NCCL inventory test, the multi-rank PCIe microbenchmark and the
# bench line at N GPUs (strong scaling of ONE EC60to30 mesh; RRS18to6 slab per GPU at N = 8).
# Usage: N=2 TAG=r02 scripts/gpu_multi.sh
set -u
N=${N:-2}; TAG=${TAG:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests/test_multi_gpu_nccl.py -q -m gpu -rP > gpurun_out/pytest_nccl_${N}gpu_$TAG.log 2>&1; echo "pytest nccl exit $?"; tail -3 gpurun_out/pytest_nccl_${N}gpu_$TAG.log
fi
if [ "${SKIP_PCIE:-0}" != "1" ]; then
  $TR --master-port 29601 scripts/micro/pcie_multi.py > gpurun_out/pcie_multi_${N}gpu_$TAG.json 2> gpurun_out/pcie_multi_${N}gpu_$TAG.err; echo "pcie exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/pcie_multi_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print({k: v for k, v in d.items() if k.startswith("aggregate")}, "numa", d.get("numa_nodes"), "cpus", d.get("cpus"))
    print([ (r["rank"], r["gpu_numa_node"], r.get("default")) for r in d["ranks"]][:2])
except Exception as e:
    print("pcie: no result", e)
PY
fi
$TR --master-port 29602 bench.py --gpus $N ${BENCH_ARGS:-} > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], d["scaling"], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "graph", d["config"]["cuda_graph"],
          "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"))
    print("  inventory", d["inventory_check"])
    print("  kernels", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_launch"].items()})
    r = (d.get("secondary") or {}).get("rrs18to6_slab")
    if r: print("  rrs", {k: r.get(k) for k in ("ms_per_step", "value", "roofline_step", "inventory_check", "skipped")})
except Exception as e:
    print("bench: no result:", e)
PY
tail -3 gpurun_out/bench_${N}gpu_$TAG.err
