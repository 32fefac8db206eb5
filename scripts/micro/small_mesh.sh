#!/bin/bash
# Where does the step time go when a GPU owns only 29 395 columns (the per-GPU slab of the EC60to30
# mesh at 8 GPUs)?  One GPU, no NCCL: the default step, without the inventory, without the CUDA
# graph, with the carbonate join inside BGC_SourceSink.
mkdir -p gpurun_out
for args in "" "--no-inventory" "--no-graph" "--strict-join"; do
  tag=$(echo "default$args" | tr -d ' -')
  python bench.py --columns ${COLS:-29395} --steps 200 --no-e2e --no-cpu --no-secondary $args > gpurun_out/small_$tag.json 2>gpurun_out/small_$tag.err
  python -c "
import json; d=json.loads(open('gpurun_out/small_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), d['gpu_launches'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
done
