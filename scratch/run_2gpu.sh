mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --mode sharded > gpurun_out/sharded_p2p_2gpu.json 2> gpurun_out/sharded_p2p_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/sharded_p2p_2gpu.json)"; python -c "
import json; d=json.loads(open('gpurun_out/sharded_p2p_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('phases_ms_rank0'))"
timeout 300 $TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/sharded_p2p_2gpu.json').read().strip().splitlines()[-1])
PY
