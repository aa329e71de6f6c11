#!/usr/bin/env python
"""Device-resident timing sweep over the BASELINE parity configs (configs 1, 3 and the NTT sweep 5):
encode at N = 2^7 .. 2^19 and forward NTT at N = 2^10 .. 2^20, 4096-byte blocks, CUDA events around `reps` back-to-back
calls.  Arrays below ~100 MB stay in the 126 MB L2 between calls (marked); the small orders are launch-latency bound.
Prints one JSON line per point."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import fastecc_b200 as fe

P = 0xFFF00001


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    fe.init(0)
    S = 1024
    only_extras = "--extras" in sys.argv
    for op, Ls in (() if only_extras else (("encode", range(7, 20)), ("ntt", range(10, 21)))):
        for L in Ls:
            N = 1 << L
            x = (torch.arange(N * S, device="cuda", dtype=torch.int64) % P).to(torch.int32).view(N, S)
            reps = max(5, min(200, int(2e9 // (N * S * 4))))
            ms = timed((lambda: fe.rs_encode_dev(x)) if op == "encode" else (lambda: fe.ntt_dev(x, False)), reps)
            nbytes = N * S * 4
            print(json.dumps({"op": op, "log_n": L, "block_bytes": 4096, "array_MiB": nbytes / 2**20, "ms": round(ms, 5), "reps": reps,
                              "GBps_read_plus_write": round(2 * nbytes / ms / 1e6, 1),
                              "MiBps_reference_convention": round((2 if op == "encode" else 1) * nbytes / 2**20 / ms * 1e3),
                              "l2_resident": nbytes < 100e6}), flush=True)
            del x
            torch.cuda.empty_cache()
    n, W = 1 << 19, 1024                                      # byte <-> GF(p) recoding, 2 GiB of random bytes (22 % of the blocks need recoding)
    raw = torch.randint(0, 256, (n, 4 * W), dtype=torch.uint8, device="cuda")
    words = torch.zeros((n, W + 4), dtype=torch.int32, device="cuda")
    ms = timed(lambda: fe.bytes_to_gfp_dev(raw, words), 10)
    print(json.dumps({"op": "bytes_to_gfp", "blocks": n, "block_bytes": 4 * W, "ms": round(ms, 5), "GBps_read_plus_write": round((n * 4 * W + n * 4 * (W + 1)) / ms / 1e6, 1)}), flush=True)
    back = torch.empty_like(raw)
    ms = timed(lambda: fe._check(fe.lib().fastecc_b200_gfp_to_bytes_dev(words.data_ptr(), back.data_ptr(), n, W, W + 4, torch.cuda.current_stream().cuda_stream)), 10)
    print(json.dumps({"op": "gfp_to_bytes", "blocks": n, "block_bytes": 4 * W, "ms": round(ms, 5), "GBps_read_plus_write": round((n * 4 * W + n * 4 * (W + 1)) / ms / 1e6, 1),
                      "roundtrip_ok": bool((back == raw).all())}), flush=True)
    del raw, words, back
    torch.cuda.empty_cache()
    from fastecc_b200 import decoder                          # erasure decoding at the headline order: 2^19 data + 2^19 parity blocks of 4 KiB
    N2 = 1 << 20
    code = (torch.arange(N2 * S, device="cuda", dtype=torch.int64) % P).to(torch.int32).view(N2, S)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    erased = torch.sort(torch.randperm(N2, device="cuda", generator=g)[:N2 // 2]).values.cpu().tolist()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    pat = decoder.ErasurePattern(N2, erased, code.device)
    torch.cuda.synchronize()
    t_pat = (time.perf_counter() - t0) * 1e3
    ms = timed(lambda: pat.recover(code), 3)                 # (timing only: the workspace is not a code word after the first call)
    print(json.dumps({"op": "decode", "log_n": 19, "erased": N2 // 2, "block_bytes": 4096, "pattern_setup_ms_wall": round(t_pat, 1), "recover_ms": round(ms, 4),
                      "GBps_codeword_bytes": round(N2 * S * 4 / ms / 1e6, 1)}), flush=True)
    del code, pat
    torch.cuda.empty_cache()
    N = 1 << 19
    for K in (1, 2, 3, 6):
        x = (torch.arange(N * S, device="cuda", dtype=torch.int64) % P).to(torch.int32).view(N, S)
        ms = timed(lambda: fe.rs_encode_asym_dev(x, N >> K), 5)
        print(json.dumps({"op": "encode_asym", "log_n": 19, "log_m": 19 - K, "block_bytes": 4096, "ms": round(ms, 5),
                          "GBps_data_plus_parity": round((N + (N >> K)) * S * 4 / ms / 1e6, 1)}), flush=True)
        del x
        torch.cuda.empty_cache()
    return 0


if __name__ == "__main__":
    sys.exit(main())
