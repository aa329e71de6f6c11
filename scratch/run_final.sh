mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
cat gpurun_out/bench_r01_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launch_list.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -c 3 -o gpurun_out/passes_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
