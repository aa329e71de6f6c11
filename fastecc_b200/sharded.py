"""One Reed-Solomon encode sharded over G GPUs (BASELINE config 4): local passes + two all-to-alls of whole blocks.

Blocks are dealt cyclically (global block i = l*G + rank is local block l) for the data going in and the parity
coming out.  N = N1*N2 as in csrc/plan.h; with N2 % G == 0 the row sets of pass A (fixed n2) and pass D (fixed j2)
are local to rank n2 % G / j2 % G, and those of the fused pass BC (fixed k1) to rank k1 % G, so the transpose
between them (TransposeMatrix on block pointers in the reference, ntt.cpp:322-341,415,433,445) becomes an all-to-all
of 4 KiB blocks over NVLink.  `run_pass(tensor, which)` executes one local pass in place: on the GPUs
fastecc_b200_rs_encode_shard_pass, in the CPU tests the emulation of the same kernel."""
from __future__ import annotations

from typing import Callable


def geometry(N: int, G: int):
    LN = N.bit_length() - 1
    if N != 1 << LN or LN < 11 or LN > 19 or G < 2 or G & (G - 1):
        raise ValueError("sharded encode needs N = 2^11..2^19 and a power-of-two number of ranks")
    L1 = min(9, (LN + 1) // 2)                       # csrc/plan.h split_l1()
    L1 = max(L1, LN - 10, 5)
    N1, N2 = 1 << L1, 1 << (LN - L1)
    if N2 % G:
        raise ValueError("N2 must be a multiple of the number of ranks")
    return N1, N2


def _all_to_all(send, group):
    import torch
    import torch.distributed as dist
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def exchange_after_a(x, N: int, G: int, group=None):
    """local [k1][n2'] (k1 = k1'*G + dest)  ->  on rank dest: [k1'][n2] with n2 = n2'*G + src."""
    N1, N2 = geometry(N, G)
    S = x.shape[1]
    send = x.view(N1 // G, G, N2 // G, S).permute(1, 0, 2, 3).contiguous()        # [dest][k1'][n2']
    recv = _all_to_all(send, group)                                               # [src][k1'][n2']
    return recv.permute(1, 2, 0, 3).contiguous().view(N // G, S)                  # [k1'][n2'][src]


def exchange_after_bc(x, N: int, G: int, group=None):
    """local [k1'][j2] (j2 = j2'*G + dest)  ->  on rank dest: [k1][j2'] with k1 = k1'*G + src."""
    N1, N2 = geometry(N, G)
    S = x.shape[1]
    send = x.view(N1 // G, N2 // G, G, S).permute(2, 0, 1, 3).contiguous()        # [dest][k1'][j2']
    recv = _all_to_all(send, group)                                               # [src][k1'][j2']
    return recv.permute(1, 0, 2, 3).contiguous().view(N // G, S)                  # [k1'][src][j2']


def rs_encode_sharded(x_local, N: int, G: int, run_pass: Callable, group=None):
    """x_local: this rank's N/G blocks (2-D, contiguous).  Returns this rank's N/G parity blocks (a new tensor)."""
    geometry(N, G)
    if x_local.dim() != 2 or x_local.shape[0] != N // G or not x_local.is_contiguous():
        raise ValueError("x_local must be a contiguous [N/G, SIZE] tensor")
    run_pass(x_local, 0)
    x = exchange_after_a(x_local, N, G, group)
    run_pass(x, 1)
    x = exchange_after_bc(x, N, G, group)
    run_pass(x, 2)
    return x


def gpu_pass_runner(N: int, G: int, rank: int):
    """run_pass for CUDA tensors: the C ABI on the current torch stream."""
    import torch
    import fastecc_b200 as fe

    def run(t, which: int):
        if not (t.is_cuda and t.is_contiguous() and t.shape[1] % 4 == 0):
            raise ValueError("sharded encode needs contiguous CUDA blocks with SIZE % 4 == 0")
        rc = fe.lib().fastecc_b200_rs_encode_shard_pass(t.data_ptr(), N, G, rank, t.shape[1], t.shape[1], which,
                                                        torch.cuda.current_stream(t.device).cuda_stream)
        fe._check(rc)
    return run
