mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
FASTECC_B200_KERNEL=cta timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "many_tiles" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "many_tiles" 2>&1 | tail -8
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ntt_pass_kernel -c 1 -o gpurun_out/bc_full -f $B --steps 1 --warmup 1 > gpurun_out/ncu_bc.log 2>&1
ls -la gpurun_out/
