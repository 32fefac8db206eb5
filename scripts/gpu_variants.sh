#!/bin/bash
# A/B of the column-sweep launch shapes on one B200 (tuning aid; not part of the product path).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-100 1 2 3 4 5 6}; do
  BGC_ECO_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_variant_$v.json 2> gpurun_out/bench_variant_$v.err
  echo "variant $v exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_variant_$v.json"))
    print("variant $v: step %.3f ms value %.4g  kernels %s" % (d["ms_per_step"], d["value"], {k: round(x,3) for k,x in d["roofline"]["kernel_ms_per_launch"].items()}))
except Exception as e:
    print("variant $v: no result", e)
PY
done
