mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --mode sharded > gpurun_out/sharded_p2p_2gpu.json 2> gpurun_out/sharded_p2p_2gpu.err
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --mode sharded-a2a > gpurun_out/sharded_a2a_2gpu.json 2> gpurun_out/sharded_a2a_2gpu.err
cut -c1-400 gpurun_out/sharded_p2p_2gpu.json; tail -3 gpurun_out/sharded_p2p_2gpu.err
cut -c1-400 gpurun_out/sharded_a2a_2gpu.json
