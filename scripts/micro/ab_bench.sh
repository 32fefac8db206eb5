#!/bin/bash
# A/B of kernel variants on the EC60to30 step, device resident: "NAME=VALUE[,NAME=VALUE] ..." sets of environment
# variables, one bench line each (step time and per-kernel times).  COLS=... LEVELS=... for another mesh shape.
mkdir -p gpurun_out
for cfg in "$@"; do
  env $(echo "$cfg" | tr ',' ' ') python bench.py --columns ${COLS:-235160} --levels ${LEVELS:-60} --steps ${STEPS:-30} --warmup 3 --no-e2e --no-cpu --no-secondary \
      > gpurun_out/ab.json 2> gpurun_out/ab.err || { tail -5 gpurun_out/ab.err; continue; }
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print('%-40s %.4f ms/step  no-shortcut %s  %s' % ('$cfg', d['ms_per_step'], round(d.get('without_zero_biomass_shortcut',{}).get('ms_per_step',0),4), {k: round(v,4) for k,v in d['roofline']['kernel_ms_per_launch'].items()}))"
done
