env | grep -E "OMP|KMP|GOMP" 
for m in plain torch bind; do timeout 120 python scratch/refarm_probe.py $m 2>&1 | tail -8; done
timeout 300 python tools/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; cut -c1-220 gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err
