#!/bin/bash
# N-GPU check, for `gpurun --gpus N --timeout 1200 -- 'bash tools/gpu/multi_gpu_check.sh N'` (charged N x): the sharded-encode
# GPU tests (fused peer stores and NCCL all-to-all), then three bench lines: one transform sharded with fused stores, the
# same with NCCL all-to-alls, and independent stripes (the default mode the driver's scaling run uses).
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --mode sharded     > gpurun_out/sharded_p2p_${N}gpu.json 2> gpurun_out/sharded_p2p_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --mode sharded-a2a > gpurun_out/sharded_a2a_${N}gpu.json 2> gpurun_out/sharded_a2a_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/stripes_${N}gpu.json 2> gpurun_out/stripes_${N}gpu.err
for f in sharded_p2p sharded_a2a stripes; do echo "$f: $(wc -l < gpurun_out/${f}_${N}gpu.json) stdout line(s)"; cut -c1-260 gpurun_out/${f}_${N}gpu.json; done
