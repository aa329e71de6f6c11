set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/split_v8.json 2> gpurun_out/split_v8.err
timeout 300 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__waves_per_multiprocessor --clock-control none -k regex:ntt_pass -c 9 --csv --log-file gpurun_out/split_v8_passes.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
FASTECC_B200_KERNEL=warp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/split_warp.json 2> gpurun_out/split_warp.err
FASTECC_B200_SPLIT=hi timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/splithi_v8.json 2> gpurun_out/splithi_v8.err
cat gpurun_out/split_v8.json gpurun_out/split_warp.json gpurun_out/splithi_v8.json | cut -c1-200
