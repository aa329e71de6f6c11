#!/bin/bash
# One-GPU A/B of build / environment variants: device-resident bench line (with the in-run parity hash and per-pass kernel
# times) of the default library, of the default library under each environment in ENVS, and of every scratch/lib_*.so.
# Run from the repo root under gpurun.  TESTS=1 also runs the GPU tests first; NCU=1 adds a launch list and a full capture.
mkdir -p gpurun_out
[ -n "$TESTS" ] && timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
i=0
for e in ${ENVS:-}; do i=$((i+1)); env $(echo $e | tr ',' ' ') timeout 300 $B > gpurun_out/bench_env$i.json 2> gpurun_out/bench_env$i.err; echo "env$i = $e"; done
for lib in scratch/lib_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so)
  FASTECC_B200_LIB=$PWD/$lib timeout 300 $B > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s %.3f ms/step" % (f[17:-5], d["ms_per_step"]), [(k["pass"], k["kernel"], k["ms"]) for k in d["roofline"].get("per_kernel", [])], "golden", d["parity"]["device_resident"]["golden_match"])
    except Exception as e:
        print(f, "unreadable", e, open(f[:-4] + "err").read()[-300:])
PY
if [ -n "$NCU" ]; then
  S="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ntt_pass -c 60 --csv --log-file gpurun_out/launch_list.csv $S > /dev/null 2>&1
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 6 -c 3 -o gpurun_out/passes_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
fi
ls gpurun_out | tr '\n' ' '
