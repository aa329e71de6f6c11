mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 240 $TR bench.py --gpus 8 --steps 20 --warmup 3 --mode sharded > gpurun_out/sharded_p2p_8gpu.json 2> gpurun_out/sharded_p2p_8gpu.err
timeout 240 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/stripes_8gpu.json 2> gpurun_out/stripes_8gpu.err
cut -c1-330 gpurun_out/sharded_p2p_8gpu.json; tail -2 gpurun_out/sharded_p2p_8gpu.err
cut -c1-330 gpurun_out/stripes_8gpu.json; tail -2 gpurun_out/stripes_8gpu.err
