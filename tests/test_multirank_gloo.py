"""CPU, world_size 2, gloo: the N>1 host logic of the independent-stripes mode (stripe ownership, the MAX-over-ranks timing
rule, gather of parity hashes, aggregate metric).  Each rank encodes ITS OWN stripes with the CPU oracle standing in for the GPU
call and rank 0 checks every stripe's parity hash against a single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def stripes_of(rank, world, n_stripes):
    """Round-robin ownership: stripe s belongs to rank s % world."""
    return list(range(rank, n_stripes, world))


def gather_ints(values):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def aggregate_throughput(bytes_per_stripe, stripes_per_rank, world, seconds):
    """Whole-job GB/s: all ranks' bytes over the max-over-ranks time (weak scaling: per-GPU work is fixed)."""
    return world * stripes_per_rank * bytes_per_stripe / seconds / 1e9


def _worker(rank, world, port, n_stripes, L, S, q):
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    from fastecc_b200 import multirank as mr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = ol.load_oracle()
    mine = stripes_of(rank, world, n_stripes)
    hashes = []
    for s in mine:
        a = ol.fill_B(o, 1 << L, S)
        a[0, 0] = s                                  # make every stripe distinct
        hashes.append(ol.ohash(o, ol.o_encode(o, a)))
    dist.barrier()
    t = mr.max_over_ranks(0.25 * (rank + 1))         # pretend rank r took 0.25*(r+1) s
    allh = gather_ints(hashes)
    if rank == 0:
        q.put((t, allh, aggregate_throughput(1e9, len(mine), world, t)))
    dist.destroy_process_group()


def test_two_rank_stripe_sharding_gloo():
    import torch.multiprocessing as mp
    import oracle_lib as ol
    world, n_stripes, L, S = 2, 4, 6, 5
    assert stripes_of(0, 2, 5) == [0, 2, 4] and stripes_of(1, 2, 5) == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_stripes, L, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, allh, gbps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == pytest.approx(0.5)                                   # MAX over ranks, not mean, not rank 0's
    assert gbps == pytest.approx(2 * 2 * 1e9 / 0.5 / 1e9)            # all ranks' bytes over the slowest rank's time
    o = ol.load_oracle()
    want = []
    for s in range(n_stripes):
        a = ol.fill_B(o, 1 << L, S); a[0, 0] = s
        want.append(ol.ohash(o, ol.o_encode(o, a)))
    got = {}
    for r in range(world):
        for s, h in zip(stripes_of(r, world, n_stripes), allh[r]):
            got[s] = h
    assert [got[s] for s in range(n_stripes)] == want
