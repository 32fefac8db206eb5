#!/bin/bash
# ncu capture of the column sweep on a reduced mesh (65536 columns x 60 levels), after a plain run.
set -u
mkdir -p gpurun_out
CMD="python bench.py --columns 65536 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:eco_columns -s 4 -c 1 -f -o gpurun_out/${NCU_OUT:-eco_v2} $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu.log
