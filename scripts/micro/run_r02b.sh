set -u
mkdir -p gpurun_out
cd /root/repo
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "carbonate_placement or cold_and_warm or block_width" 2>&1 | tail -5
for sh in 60 85 100 115; do BGC_CO3_SHARE=$sh python scripts/micro/concurrency_sweep.py 29396 14698; done > gpurun_out/concurrency_sweep_r02b.txt 2>&1
BGC_CO3_CONFINED=0 python scripts/micro/concurrency_sweep.py 29396 58790 235160 >> gpurun_out/concurrency_sweep_r02b.txt 2>&1
cat gpurun_out/concurrency_sweep_r02b.txt
RW_MIX_TMA=1 ./scripts/micro/rw_mix > gpurun_out/rw_mix_tma_r02.txt 2>&1; cat gpurun_out/rw_mix_tma_r02.txt
