#!/bin/bash
# One-GPU round check, meant for `gpurun --timeout 2400 -- 'bash tools/gpu/single_gpu_check.sh'` (≈ 4-5 GPU-minutes):
# GPU tests, the default bench line, the reference arm, smoke, the per-launch list and one `--set full` capture of the
# three passes of an encode (A, BC, D).  Everything lands in gpurun_out/; summarise with tools/ncu_summary.py and copy what
# should be judged into profiles/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "stdout lines: $(wc -l < gpurun_out/bench.json)"; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ntt_pass -c 60 --csv --log-file gpurun_out/launch_list.csv $B > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -c 3 -o gpurun_out/passes_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
