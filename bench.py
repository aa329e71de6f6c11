#!/usr/bin/env python
"""bench.py -- RS encode GB/s at (n,k)=(2^20,2^19), 4 KiB blocks (BASELINE.json metric), on N B200s.

Own arm (default): one process per GPU (torchrun for N>1).  A step = one full encode (RS.cpp:41-63) of
N=2^19 data blocks x 4096 B resident in HBM -> 2^19 parity blocks, through the C ABI
(fastecc_b200_rs_encode_dev).  Throughput convention is the reference's: bytes = 2*N*SIZE*4 per encode
(RS.cpp:38).  With N>1 GPUs every rank encodes its own independent stripe of 2^19 blocks (weak scaling, no
data-path collective).

`--impl reference`: times the UNMODIFIED reference CPU encoder (oracle/_ref, compiled from /root/reference by
oracle/Makefile: AVX2 + OpenMP build, all host threads) on the same config; rank 0 only.

One JSON line on stdout (rank 0).  See DESIGN.md section 7 for how every field is measured.
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
P = 0xFFF00001
METRIC = "rs_encode_GBps_n2^20_k2^19_4KiB_blocks"


def profiled_traffic():
    """DRAM bytes (read + write) per ntt_pass_kernel launch from the committed `ncu --set full` capture of this workload
    (profiles/*_passes_A_BC_D.csv, highest round / version in the name), averaged over the three passes of an encode; None if absent."""
    import csv
    import glob
    import re

    def version(path):                      # r<round>_ncu_v<kernel version>_passes_A_BC_D.csv: newest round, then newest version
        m = re.search(r"r(\d+)_ncu_v(\d+)_passes", os.path.basename(path))
        return (int(m.group(1)), int(m.group(2))) if m else (-1, -1)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_v*_passes_A_BC_D.csv")), key=version)
    if not files:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        h, units, data = rows[0], rows[1], rows[2:]
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in data]
        return sum(tot) / len(tot), os.path.basename(files[-1])
    except Exception:
        return None, None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=19, help="log2 of the number of data blocks (headline: 19)")
    ap.add_argument("--block-bytes", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mode", default="stripes", choices=["stripes", "sharded", "sharded-a2a"],
                    help="multi-GPU: independent stripe per GPU (default, weak scaling) or ONE transform sharded over the GPUs with two NCCL all-to-alls (strong scaling, BASELINE config 4)")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = sorted(s for s, p in zip(sm, pw) if p >= 0.6 * max(pw)) or sorted(sm)
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- reference CPU arm
def load_ref_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "libfastecc_ref.so")
    if not os.path.exists(path):
        return None
    r = ctypes.CDLL(path)
    r.ref_rs_encode.restype = None                      # RS.cpp:41-63 on a caller-supplied T** table
    r.ref_rs_encode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
    r.ref_num_threads.restype = ctypes.c_int
    r.ref_build_flavour.restype = ctypes.c_char_p
    if hasattr(r, "ref_set_num_threads"):
        r.ref_set_num_threads.argtypes = [ctypes.c_int]; r.ref_set_num_threads.restype = None
    return r


def host_thread_candidates():
    """Thread counts worth trying for the CPU reference: one per physical core and one per hardware thread of this
    process's affinity mask (the reference blocks for the caches of a core: SMT siblings can hurt it, 1.8 s vs 4.6 s
    per encode on a 64-core / 128-thread host)."""
    cpus = sorted(os.sched_getaffinity(0))
    cores = set()
    for c in cpus:
        try:
            cores.add(open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read().strip())
        except OSError:
            cores.add(str(c))
    return sorted({max(1, len(cores)), len(cpus)})


def load_oracle_port():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib.load_oracle()


def fill_index_mod_p(arr, chunk=1 << 24):
    """arr.flat[i] = i % P (the reference's fill, RS.cpp:28-29), in slices: no array-sized temporaries (eight ranks fill
    2 GiB each at the same time in the multi-GPU runs)."""
    import numpy as np
    flat = arr.reshape(-1)
    for lo in range(0, flat.size, chunk):
        hi = min(lo + chunk, flat.size)
        flat[lo:hi] = (np.arange(lo, hi, dtype=np.uint64) % P).astype(np.uint32)


def cpu_encode_runner(log_n, size_words, calibrate=True):
    """Returns (fn, kind, cores, label): fn() runs one full CPU encode of 2^log_n x size_words in place."""
    import numpy as np
    N = 1 << log_n
    buf = np.empty(N * size_words, dtype=np.uint32)
    fill_index_mod_p(buf)
    r = load_ref_lib()
    if r is not None:
        # T** data, RS.cpp:31-33; left permuted between steps like the reference leaves it
        tab = buf.ctypes.data + np.arange(N, dtype=np.uint64) * np.uint64(size_words * 4)
        fn = lambda _keep=buf: r.ref_rs_encode(tab.ctypes.data, N, size_words)      # noqa: E731
        note = ""
        cands = host_thread_candidates()
        if hasattr(r, "ref_set_num_threads") and calibrate:
            best = None
            for n in cands:                                   # give the reference its best thread count on this host
                r.ref_set_num_threads(n)
                if best is None:
                    fn()                                      # first touch / page faults are not part of the comparison
                t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
                if best is None or dt < best[1]:
                    best = (n, dt)
            r.ref_set_num_threads(best[0])
            note = "; thread count picked from %s by one timed encode each" % cands
        return fn, "reference", int(r.ref_num_threads()), \
            "unmodified FastECC templates, %s+OpenMP build (oracle/_ref)%s" % (r.ref_build_flavour().decode(), note)
    o = load_oracle_port()
    return (lambda: o.oracle_rs_encode(buf.ctypes.data, N, size_words)), "port", int(o.oracle_num_threads()), "oracle/gfp_oracle.c (plain C port, OpenMP)"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ.setdefault("OMP_WAIT_POLICY", "active")
    os.environ["OMP_NUM_THREADS"] = str(host_thread_candidates()[0])       # torchrun pins it to 1; cpu_encode_runner() then picks the better of cores / hardware threads
    size_words = args.block_bytes // 4
    N = 1 << args.log_n
    fn, kind, cores, label = cpu_encode_runner(args.log_n, size_words)
    for _ in range(max(1, min(args.warmup, 3))):
        fn()
    times = []
    budget_t0 = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
        if time.perf_counter() - budget_t0 > 240:          # keep the whole run within a few minutes
            break
    total = sum(times)
    nbytes = 2.0 * N * size_words * 4
    value = nbytes * len(times) / total / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "best_ms": 1e3 * min(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "rs_encode N=2^%d data blocks -> 2^%d parity, %d-byte blocks, GF(0xFFF00001), host memory" % (args.log_n, args.log_n, args.block_bytes),
                   "bytes_per_step": nbytes, "convention": "2*N*SIZE*4 bytes per encode (RS.cpp:38)"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": kind, "sample": "%d full encodes of the workload; %s" % (len(times), label)},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)
    return 0


def quick_cpu_baseline(args):
    """Bounded CPU sample for the own arm's cpu_baseline block: a few full encodes (about 10-30 s of CPU work at most)."""
    try:
        os.environ.setdefault("OMP_WAIT_POLICY", "active")
        os.environ["OMP_NUM_THREADS"] = str(host_thread_candidates()[0])
        size_words = args.block_bytes // 4
        fn, kind, cores, label = cpu_encode_runner(args.log_n, size_words)
        fn()
        times = []
        t_start = time.perf_counter()
        while len(times) < 5 and time.perf_counter() - t_start < 20:
            t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
        nbytes = 2.0 * (1 << args.log_n) * size_words * 4
        return {"value": nbytes / min(times) / 1e9, "unit": "GB/s", "cores": cores, "kind": kind,
                "sample": "best of %d full encodes of the same workload (mean %.1f ms); %s" % (len(times), 1e3 * sum(times) / len(times), label)}
    except Exception as e:       # never let the baseline leg break the GPU measurement
        return {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}


# ---------------------------------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's banner off stdout: rank 0 prints exactly one JSON line there
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fastecc_b200 as fe
    fe.init(local)                      # raises if the CUDA library or an sm_100 GPU is missing: no fallback

    N, S = 1 << args.log_n, args.block_bytes // 4
    dev = torch.device("cuda", local)
    if args.mode in ("sharded", "sharded-a2a") and world > 1:
        return run_sharded(args, fe, rank, world, local, dev)
    data = torch.empty((N, S), dtype=torch.int32, device=dev)
    flat = data.view(-1)
    step_elems = 1 << 26
    for lo in range(0, flat.numel(), step_elems):                      # fill A: data0[i] = i % P  (RS.cpp:28-29)
        hi = min(lo + step_elems, flat.numel())
        flat[lo:hi] = (torch.arange(lo, hi, device=dev, dtype=torch.int64) % P).to(torch.int32)
    nbytes = 2.0 * N * S * 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        fe.rs_encode_dev(data)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = fe.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        fe.rs_encode_dev(data)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = fe.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None
    from fastecc_b200 import multirank
    ms = multirank.max_over_ranks(ms, device=dev)          # a multi-GPU step takes as long as its slowest rank
    barrier()

    # ---- end-to-end through the reference-facing host call (T** table, pinned host memory, H2D + D2H inside)
    e2e = None
    if not args.no_e2e:
        hptr = fe.lib().fastecc_b200_host_alloc(N * S * 4)
        if not hptr:
            raise SystemExit("pinned allocation failed")
        harr = np.ctypeslib.as_array((ctypes.c_uint32 * (N * S)).from_address(hptr)).reshape(N, S)
        fill_index_mod_p(harr)
        fe.EncodeReedSolomon_body(harr, N, S)                           # warm-up (allocates the device staging buffer)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fe.EncodeReedSolomon_body(harr, N, S)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_s = multirank.max_over_ranks(e2e_s, device=dev)
        e2e = {"value": world * nbytes * args.e2e_steps / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": N * S * 4, "d2h_bytes_per_step": N * S * 4,
               "ms_per_step": 1e3 * e2e_s / args.e2e_steps, "api": "fastecc_b200_rs_encode(T** data, N, SIZE) on pinned host blocks"}
        del harr
        fe.lib().fastecc_b200_host_free(hptr)

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic, traffic_src = profiled_traffic() if (args.log_n == 19 and args.block_bytes == 4096) else (None, None)
        passes_per_step = launches / max(args.steps, 1)
        launch_ms = ms / max(launches, 1)
        achieved = nbytes / (launch_ms * 1e-3) / 1e9              # every pass reads and writes the whole array once
        out = {
            "metric": METRIC, "value": world * nbytes * args.steps / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "rs_encode N=2^%d data blocks -> 2^%d parity, %d-byte blocks, GF(0xFFF00001), resident in HBM" % (args.log_n, args.log_n, args.block_bytes),
                       "bytes_per_step": nbytes, "convention": "2*N*SIZE*4 bytes per encode (RS.cpp:38)", "per_gpu_buffer_bytes": N * S * 4,
                       "l2": "inputs (2 GiB per GPU) are larger than L2; no flush needed", "parallelism": "independent stripe per GPU" if world > 1 else "single GPU",
                       "passes_per_encode": passes_per_step},
            "roofline": {"bound": "hbm", "kernel": "ntt_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "note": "algorithmic bytes per launch = 2*N*SIZE*4 (one read + one write of the array per pass); avg launch = timed region / launches"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e:
            out["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = quick_cpu_baseline(args)
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_sharded(args, fe, rank, world, local, dev):
    """ONE encode of 2^log_n blocks sharded over the ranks: local passes + two all-to-alls (fastecc_b200/sharded.py)."""
    import torch
    import torch.distributed as dist
    from fastecc_b200 import multirank, sharded
    N, S = 1 << args.log_n, args.block_bytes // 4
    rows = N // world
    x = (torch.arange(rows * S, device=dev, dtype=torch.int64) * 2654435761 % P).to(torch.int32).view(rows, S)
    nbytes = 2.0 * N * S * 4
    fused = args.mode == "sharded" and sharded.p2p_supported(N, world)
    if fused:                                   # exchange fused into the kernels' stores over peer memory
        enc = sharded.P2PShardedEncoder(N, S)
        enc.x.copy_(x)

        def step(_):
            return enc.encode()
    else:                                       # local passes + two NCCL all-to-alls
        run_pass = sharded.gpu_pass_runner(N, world, rank)

        def step(t):
            return sharded.rs_encode_sharded(t, N, world, run_pass)

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        x = step(x)
    sync()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.25)
    launches0 = fe.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    for _ in range(args.steps):
        x = step(x)
    ev1.record()
    torch.cuda.synchronize()
    ms = multirank.max_over_ranks(ev0.elapsed_time(ev1), device=dev)
    launches = fe.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None
    sync()
    phases = None
    if fused:                                   # one more encode with events between the phases (rank 0's view)
        evs = []
        enc.encode(events=evs)
        torch.cuda.synchronize()
        phases = dict(zip(["pass_A", "barrier_1", "pass_BC", "barrier_2", "pass_D"], [round(evs[i].elapsed_time(evs[i + 1]), 4) for i in range(5)]))
        sync()
    if rank == 0:
        a2a_bytes = 2.0 * (world - 1) / world * (N * S * 4 / world)           # sent per GPU per encode (two all-to-alls)
        link = 770.0                                                          # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)
        t_link = a2a_bytes / (link * 1e9)
        out = {
            "metric": METRIC, "value": nbytes * args.steps / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "rs_encode N=2^%d data blocks -> 2^%d parity, %d-byte blocks, ONE transform sharded over %d GPUs" % (args.log_n, args.log_n, args.block_bytes, world),
                       "bytes_per_step": nbytes, "convention": "2*N*SIZE*4 bytes per encode (RS.cpp:38)", "parallelism": ("cyclic blocks, 3 passes; passes A and BC store every output row into its owner's HBM over NVLink (peer-mapped, CUDA IPC), 2 one-word all-reduce barriers"
                                       if fused else "cyclic blocks, 3 local passes + 2 NCCL all-to-all"),
                       "l2": "local arrays (%.0f MiB per GPU) exceed L2" % (N * S * 4 / world / 2**20)},
            "roofline": {"bound": "nvlink", "achieved": a2a_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": link, "unit": "GB/s",
                         "frac": t_link / (ms / args.steps * 1e-3), "traffic": None,
                         "note": "bytes each GPU sends to its peers per encode (2 exchanges of (G-1)/G of the local array) / step time, against the measured 770 GB/s per-direction peer bandwidth"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if phases:
            out["phases_ms_rank0"] = phases
        emit(out)
    if fused:
        enc.close()
    dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def guard_stdout():
    """Libraries write banners to stdout (torch prints "NCCL version ..." when the first communicator is created).  The
    contract is ONE JSON line there, so file descriptor 1 points at stderr until emit() prints the result."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)
    if _REAL_STDOUT is not None:
        os.dup2(2, 1)


def main():
    args = parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
