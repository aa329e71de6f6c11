import os, sys, time
sys.path.insert(0, "/root/repo")
mode = sys.argv[1]
os.environ.setdefault("OMP_WAIT_POLICY", "active")
print(mode, "affinity", len(os.sched_getaffinity(0)), "OMP_NUM_THREADS env", os.environ.get("OMP_NUM_THREADS"), "cpu_count", os.cpu_count(), flush=True)
os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
if mode == "torch":
    import torch
    print("torch threads", torch.get_num_threads(), "affinity now", len(os.sched_getaffinity(0)), flush=True)
if mode == "bind":
    os.environ["OMP_PROC_BIND"] = "spread"; os.environ["OMP_PLACES"] = "threads"
import bench
fn, kind, cores, label = bench.cpu_encode_runner(19, 1024)
print(kind, cores, label, flush=True)
for i in range(4):
    t0 = time.perf_counter(); fn(); print("  %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
