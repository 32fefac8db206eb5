#!/bin/bash
# quick GPU check: parity tests + device-resident bench line (no e2e / cpu legs); VARIANTS="eco:dms ..." optional
set -u
mkdir -p gpurun_out
TAG=${TAG:-quick}
python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAIL:-6}
for v in ${VARIANTS:-0:0}; do
  BGC_ECO_VARIANT=${v%%:*} BGC_DMS_VARIANT=${v##*:} python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err || tail -5 gpurun_out/bench_${TAG}_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_$v.json')); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['roofline']['kernel_ms_per_launch'].items()})"
done
