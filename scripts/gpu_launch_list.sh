#!/bin/bash
# per-launch device times of one bench step (ncu --metrics gpu__time_duration.sum), after a plain run
set -u
mkdir -p gpurun_out
CMD="python bench.py --columns ${COLS:-235160} --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:eco_columns|co3_cells|dms_cells|dms_columns|dms_surface|zsat_columns|macros_cells|surface_fluxes|inventory_|co2calc_points|transpose" -s ${SKIP:-40} -c ${COUNT:-20} --csv --log-file gpurun_out/${OUT:-launches}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
