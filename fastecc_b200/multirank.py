"""Host-side rule of the multi-GPU measurements (one process per GPU, torch.distributed for the plumbing): a step takes as long
as its slowest rank.  Backend-agnostic ("nccl" on the GPUs, "gloo" in the CPU tests)."""
from __future__ import annotations


def max_over_ranks(value: float, device=None) -> float:
    """MAX reduction of a per-rank time (or any scalar) over the default process group; identity for a single process."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
