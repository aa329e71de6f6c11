mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
echo "stdout lines: $(wc -l < gpurun_out/bench_r01_final.json)"; cut -c1-250 gpurun_out/bench_r01_final.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/bench_r01_reference.json; cut -c1-250 gpurun_out/bench_r01_reference.json
