#!/bin/bash
# One GPU-box visit: parity tests, the default bench line, the reference arm, the per-launch
# device-time list of one step and one `ncu --set full` capture per hot kernel (full EC60to30
# mesh).  Every ncu pass runs only after the same command exited 0 without ncu.
# Usage: TAG=r01_v4 scripts/gpu_round.sh   (outputs under gpurun_out/)
set -u
TAG=${TAG:-r01}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests -m gpu ${PYTEST_ARGS:--x} -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
  tail -3 gpurun_out/pytest_gpu_$TAG.log
fi
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"
BGC_BENCH_WRITE_INVENTORY=${WRITE_INV:-} python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:eco_columns|co3_cells|dms_cells|dms_columns|dms_surface|macros_cells|surface_fluxes|zsat_columns|inventory_" -s ${SKIP:-40} -c ${COUNT:-24} --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
if [ "${SKIP_FULL:-0}" != "1" ]; then
  ncu --set full --clock-control none --import-source on \
      -k "regex:${KERNELS:-eco_columns|co3_cells|dms_cells|macros_cells|zsat_columns|surface_fluxes}" -s ${FSKIP:-16} -c ${FCOUNT:-4} -f \
      -o gpurun_out/ncu_full_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; grep -c "==PROF== Profiling" gpurun_out/ncu_full.log
fi
python - <<PY
import json
for f in ("gpurun_out/bench_$TAG.json", "gpurun_out/bench_ref_$TAG.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"],
              "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        if "roofline" in d:
            print("  kernels", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_launch"].items()})
    except Exception as e:
        print(f, "no result:", e)
PY
# optional extras (each a few seconds of box time): FUZZ=300 runs the differential fuzzer against the CUDA path,
# MICRO=1 the read/write-mix micro-benchmark (scripts/micro/rw_mix must have been built: see its header)
if [ -n "${FUZZ:-}" ]; then
  python scripts/fuzz_gpu_vs_reference.py $FUZZ ${FUZZ_SEED0:-0} > gpurun_out/fuzz_gpu_$TAG.txt 2>&1; echo "fuzz exit $?"; head -3 gpurun_out/fuzz_gpu_$TAG.txt
  BGC_B200_FLAVOUR=strict python scripts/fuzz_gpu_vs_reference.py $FUZZ ${FUZZ_SEED0:-0} > gpurun_out/fuzz_gpu_strict_$TAG.txt 2>&1; echo "fuzz (strict flavour) exit $?"; head -3 gpurun_out/fuzz_gpu_strict_$TAG.txt
fi
if [ "${MPAS_SWEEP:-0}" = "1" ]; then
  python scripts/micro/mpas_sweep.py > gpurun_out/mpas_sweep_$TAG.txt 2>&1; cat gpurun_out/mpas_sweep_$TAG.txt
fi
if [ "${MICRO:-0}" = "1" ] && [ -x scripts/micro/rw_mix ]; then
  scripts/micro/rw_mix > gpurun_out/rw_mix_$TAG.txt 2>&1; RW_MIX_STORES=1 scripts/micro/rw_mix > gpurun_out/rw_mix_stores_$TAG.txt 2>&1; tail -4 gpurun_out/rw_mix_stores_$TAG.txt
fi
