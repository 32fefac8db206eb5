#!/bin/bash
# round-2 visit B: parity with the pipelined DMS column kernel forced, A/B timings on EC60to30, the RRS18to6 slab
# and the two mixed shapes (which of columns / levels slows the tile kernel down on the RRS slab)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/vb_pytest.log 2>&1; echo "pytest default exit $?"; tail -2 gpurun_out/vb_pytest.log
BGC_DMS_VARIANT=2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/vb_pytest_2.log 2>&1; echo "pytest dms=2 exit $?"; tail -2 gpurun_out/vb_pytest_2.log
BGC_DMS_VARIANT=2 BGC_DMS_COLUMN_THREADS=128 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/vb_pytest_2_128.log 2>&1; echo "pytest dms=2 threads=128 exit $?"; tail -2 gpurun_out/vb_pytest_2_128.log
STEPS=20 scripts/micro/ab_bench.sh X=0 BGC_DMS_VARIANT=2 BGC_DMS_VARIANT=2,BGC_DMS_COLUMN_THREADS=128 BGC_DMS_VARIANT=3 2>&1 | tee gpurun_out/vb_ab_ec.txt
COLS=461654 LEVELS=80 STEPS=10 scripts/micro/ab_bench.sh X=0 BGC_DMS_COLUMN_THREADS=128 BGC_DMS_VARIANT=1 2>&1 | tee gpurun_out/vb_ab_rrs.txt
COLS=235160 LEVELS=80 STEPS=10 scripts/micro/ab_bench.sh BGC_DMS_VARIANT=1 BGC_DMS_VARIANT=2 2>&1 | tee gpurun_out/vb_ab_ec80.txt
COLS=461654 LEVELS=60 STEPS=10 scripts/micro/ab_bench.sh BGC_DMS_VARIANT=1 BGC_DMS_VARIANT=2 2>&1 | tee gpurun_out/vb_ab_rrs60.txt
