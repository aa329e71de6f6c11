"""Host-side logic of the multi-GPU path (one process per GPU, torch.distributed for the plumbing).

RS encoding shards by independent stripes (a stripe = N data blocks -> N parity blocks): rank r encodes the stripes
`stripes_of(r, world, n)` with no data-path collective; the only collectives are the barrier around the timed
region, a MAX reduction of the device time, and an optional gather of per-stripe parity hashes for verification.
Backend-agnostic ("nccl" on the GPUs, "gloo" in the CPU tests)."""
from __future__ import annotations

import os
from typing import List, Sequence


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def stripes_of(rank: int, world: int, n_stripes: int) -> List[int]:
    """Round-robin ownership: stripe s belongs to rank s % world (every rank gets floor or ceil of n/world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_stripes, world))


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_ints(values: Sequence[int], device=None) -> List[List[int]]:
    """All-gather a short list of integers (e.g. parity hashes of the stripes a rank owns); same length on every rank."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(values)]
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().tolist() for o in out]


def aggregate_throughput(bytes_per_stripe: float, stripes_per_rank: int, world: int, seconds: float) -> float:
    """Whole-job GB/s: all ranks' bytes over the max-over-ranks time (weak scaling: per-GPU work is fixed)."""
    return world * stripes_per_rank * bytes_per_stripe / seconds / 1e9
