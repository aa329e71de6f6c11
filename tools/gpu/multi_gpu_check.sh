#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu/multi_gpu_check.sh N': the multi-GPU tests (log kept) and the bench line of every mode.
# Default mode = ONE encode sharded over the N GPUs with the exchange fused into the kernels' stores (what the driver's
# scaling run records), then the NCCL all-to-all variant of the same decomposition and the stripes-only mode.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_gpu_sharded_${N}gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu_sharded_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3                    > gpurun_out/bench_sharded_p2p_${N}gpu.json 2> gpurun_out/bench_sharded_p2p_${N}gpu.err
[ -n "$QUICK" ] || timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --mode sharded-a2a --no-stripes > gpurun_out/bench_sharded_a2a_${N}gpu.json 2> gpurun_out/bench_sharded_a2a_${N}gpu.err
[ -n "$QUICK" ] || timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --mode stripes > gpurun_out/bench_stripes_${N}gpu.json 2> gpurun_out/bench_stripes_${N}gpu.err
[ -n "$QUICK" ] || timeout 600 $TR tools/sweep_sharded.py > gpurun_out/sweep_sharded_${N}gpu.jsonl 2> gpurun_out/sweep_sharded_${N}gpu.err
[ -e gpurun_out/sweep_sharded_${N}gpu.jsonl ] && cut -c1-200 gpurun_out/sweep_sharded_${N}gpu.jsonl
for f in sharded_p2p sharded_a2a stripes; do [ -e gpurun_out/bench_${f}_${N}gpu.json ] || continue; echo "$f: $(wc -l < gpurun_out/bench_${f}_${N}gpu.json) stdout line(s)"; cut -c1-2600 gpurun_out/bench_${f}_${N}gpu.json; tail -3 gpurun_out/bench_${f}_${N}gpu.err | cut -c1-400; done
