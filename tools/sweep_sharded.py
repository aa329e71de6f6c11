#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: ONE forward NTT of 2^L blocks x 4096 B sharded over the ranks (fastecc_b200/sharded.py
P2PShardedEncoder.ntt: pass A' stores into the owners' HBM over NVLink, pass B' local), L = 11 .. 20, and ONE sharded encode,
L = 12 .. 19.  Launch under torchrun (one process per GPU); CUDA events, max over ranks; rank 0 prints one JSON line per point.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_sharded.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import fastecc_b200 as fe
from fastecc_b200 import multirank, sharded

P = 0xFFF00001


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fe.init(local)
    dev = torch.device("cuda", local)
    S = 1024

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    for op, Ls in (("ntt", range(11, 21)), ("encode", range(12, 20))):
        for L in Ls:
            N = 1 << L
            ok = sharded.p2p_ntt_supported(N, world) if op == "ntt" else sharded.p2p_supported(N, world)
            if not ok:
                if rank == 0:
                    print(json.dumps({"op": op, "log_n": L, "n_gpus": world, "skipped": "first-pass tiles cannot be split over %d ranks" % world}), flush=True)
                continue
            enc = sharded.P2PShardedEncoder(N, S)
            rows = N // world
            enc.x.copy_(((torch.arange(rows * S, device=dev, dtype=torch.int64) * 2654435761) % P).to(torch.int32).view(rows, S))
            fn = (lambda: enc.ntt(False)) if op == "ntt" else enc.encode
            reps = max(5, min(100, int(1e9 // (N * S * 4 // world))))
            for _ in range(3):
                fn()
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = multirank.max_over_ranks(e0.elapsed_time(e1), device=dev) / reps
            sync()
            if rank == 0:
                nbytes = N * S * 4
                print(json.dumps({"op": op, "log_n": L, "n_gpus": world, "block_bytes": 4096, "array_MiB_total": nbytes / 2**20, "ms": round(ms, 5), "reps": reps,
                                  "GBps_read_plus_write": round(2 * nbytes / ms / 1e6, 1),
                                  "MiBps_reference_convention": round((2 if op == "encode" else 1) * nbytes / 2**20 / ms * 1e3)}), flush=True)
            enc.close()
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
