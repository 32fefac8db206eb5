#!/bin/bash
# ncu capture of selected kernels (regex in $1) on a reduced mesh, after a plain run.
set -u
mkdir -p gpurun_out
CMD="python bench.py --columns ${COLS:-65536} --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${SKIP:-8} -c ${COUNT:-2} -f -o gpurun_out/${NCU_OUT:-kernels} $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; grep -c "==PROF== Profiling" gpurun_out/ncu.log
