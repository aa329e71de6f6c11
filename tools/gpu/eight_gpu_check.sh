#!/bin/bash
# gpurun --gpus 8 -- 'bash tools/gpu/eight_gpu_check.sh': the short 8-GPU list (box time is charged 8x): the sharded GPU tests at
# G = 8 (full-size golden hashes), the default bench line (sharded headline + e2e + stripes block) and the sharded sweep.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_gpu_sharded_${N}gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu_sharded_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_sharded_p2p_${N}gpu.json 2> gpurun_out/bench_sharded_p2p_${N}gpu.err
cut -c1-3000 gpurun_out/bench_sharded_p2p_${N}gpu.json; tail -2 gpurun_out/bench_sharded_p2p_${N}gpu.err | cut -c1-300
timeout 600 $TR tools/sweep_sharded.py > gpurun_out/sweep_sharded_${N}gpu.jsonl 2> gpurun_out/sweep_sharded_${N}gpu.err
cut -c1-200 gpurun_out/sweep_sharded_${N}gpu.jsonl
