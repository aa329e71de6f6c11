"""Multi-GPU (>= 2 visible B200s): ONE encode sharded over the GPUs -- with the exchange fused into the kernels' stores
over peer memory (P2PShardedEncoder) and with NCCL all-to-alls (rs_encode_sharded) -- must equal the oracle bit for bit.  Skipped on a single-GPU box (the CPU gloo test covers the logic there)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker_p2p(rank, world, port, L, S, q):
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    import fastecc_b200 as fe
    from fastecc_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fe.init(rank)
    N = 1 << L
    o = ol.load_oracle()
    full = ol.fill_B(o, N, S)
    enc = sharded.P2PShardedEncoder(N, S)
    ok = True
    for rep in range(2):                                                   # twice: buffers and barriers are reused
        enc.x.copy_(torch.from_numpy(np.ascontiguousarray(full[rank::world]).view(np.int32)))
        out = enc.encode().clone()
        gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        dist.gather(out, gathered, dst=0)
        if rank == 0:
            par = np.empty((N, S), dtype=np.uint32)
            for r in range(world):
                par[r::world] = gathered[r].cpu().numpy().view(np.uint32)
            ok = ok and bool(np.array_equal(par, ol.o_encode(o, full)))
    enc.close()
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


def _worker_p2p_headline(rank, world, port, L, S, q):
    """BASELINE config 4 at full size: fill A (data0[i] = i % P, RS.cpp:28-29) dealt cyclically, one sharded encode, the
    gathered parity must hash (main.cpp:203-212) to the golden value of the unmodified reference; the same through
    encode_host (column-chunked H2D / passes / D2H pipeline) from a host shard."""
    import json
    import torch
    import torch.distributed as dist
    import fastecc_b200 as fe
    from fastecc_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fe.init(rank)
    N = 1 << L
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "survey_8c.json")))
    want = [g[3] for g in golden["encode_fillA"] if g[0] == L and g[1] == S][0]
    rows = N // world
    idx = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(world) + np.uint64(rank)) * np.uint64(S) + np.arange(S, dtype=np.uint64)[None, :]
    shard = (idx % np.uint64(0xFFF00001)).astype(np.uint32)
    del idx
    enc = sharded.P2PShardedEncoder(N, S)
    enc.x.copy_(torch.from_numpy(shard.view(np.int32)))
    out = enc.encode().clone()
    host = shard.copy()
    enc.encode_host(host)                                                  # pageable host memory is allowed, pinned is faster
    same = bool(np.array_equal(host, out.cpu().numpy().view(np.uint32)))
    gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, gathered, dst=0)
    flags = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        par = np.empty((N, S), dtype=np.uint32)
        for r in range(world):
            par[r::world] = gathered[r].cpu().numpy().view(np.uint32)
        q.put(bool(fe.reference_hash(par) == want and int(flags.item()) == 1 and int(par.max()) < 0xFFF00001))
    enc.close()
    dist.destroy_process_group()


def _worker_p2p_ntt(rank, world, port, L, S, q):
    """Sharded standalone transform: small orders against the oracle word for word; L >= 19 against the reference's golden
    "ntt n L 4096" hashes (fill A), plus inverse(forward(x)) == N * x."""
    import json
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    import fastecc_b200 as fe
    from fastecc_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fe.init(rank)
    N, P = 1 << L, 0xFFF00001
    rows = N // world
    enc = sharded.P2PShardedEncoder(N, S)
    ok = True
    if L < 19:
        o = ol.load_oracle()
        full = ol.fill_B(o, N, S)
        for inverse in (False, True):
            enc.x.copy_(torch.from_numpy(np.ascontiguousarray(full[rank::world]).view(np.int32)))
            out = enc.ntt(inverse).clone()
            gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
            dist.gather(out, gathered, dst=0)
            if rank == 0:
                res = np.empty((N, S), dtype=np.uint32)
                for r in range(world):
                    res[r::world] = gathered[r].cpu().numpy().view(np.uint32)
                ok = ok and bool(np.array_equal(res, ol.o_ntt(o, full, inverse)))
    else:
        golden = json.load(open(os.path.join(ROOT, "tests", "golden", "survey_8c.json")))["ntt_fillA_4096B"][str(L)]
        idx = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(world) + np.uint64(rank)) * np.uint64(S) + np.arange(S, dtype=np.uint64)[None, :]
        shard = torch.from_numpy((idx % np.uint64(P)).astype(np.uint32).view(np.int32)).cuda()
        del idx
        enc.x.copy_(shard)
        out = enc.ntt(False).clone()
        enc.ntt(True)
        want = (shard.long() & 0xFFFFFFFF) * N % P                         # unnormalised inverse: N * x
        same = bool(((enc.x.long() & 0xFFFFFFFF) == want).all())
        del want
        gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        dist.gather(out, gathered, dst=0)
        flags = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            res = np.empty((N, S), dtype=np.uint32)
            for r in range(world):
                res[r::world] = gathered[r].cpu().numpy().view(np.uint32)
            ok = bool(fe.reference_hash(res) == golden[1] and int(flags.item()) == 1)
    enc.close()
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


def _worker(rank, world, port, L, S, q):
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    import fastecc_b200 as fe
    from fastecc_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fe.init(rank)
    N = 1 << L
    o = ol.load_oracle()
    full = ol.fill_B(o, N, S)
    local = torch.from_numpy(np.ascontiguousarray(full[rank::world]).view(np.int32)).cuda()
    out = sharded.rs_encode_sharded(local, N, world, sharded.gpu_pass_runner(N, world, rank))
    gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, gathered, dst=0)
    if rank == 0:
        par = np.empty((N, S), dtype=np.uint32)
        for r in range(world):
            par[r::world] = gathered[r].cpu().numpy().view(np.uint32)
        q.put(bool(np.array_equal(par, ol.o_encode(o, full))))
    dist.destroy_process_group()


def _guarded(worker, rank, world, port, L, S, q):
    """A worker that dies must fail the test at once, not after the queue timeout (multi-GPU box time is expensive)."""
    try:
        worker(rank, world, port, L, S, q)
    except BaseException:                                  # noqa: BLE001
        import traceback
        traceback.print_exc()
        q.put(False)
        os._exit(1)


def _run(worker, L, S):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 1 << (world.bit_length() - 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_guarded, args=(worker, r, world, port, L, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = False
    try:
        ok = q.get(timeout=300)
    finally:
        for p in procs:
            p.join(timeout=120 if ok else 20)
            if p.is_alive():
                p.kill()
    assert ok
    for p in procs:
        assert p.exitcode == 0


@pytest.mark.parametrize("L,S", [(16, 64), (17, 1024)])
def test_p2p_fused_exchange_encode_on_gpus(L, S):
    _run(_worker_p2p, L, S)


@pytest.mark.parametrize("L,S", [(12, 64), (16, 256), (19, 1024), (20, 1024)])
def test_p2p_sharded_ntt_on_gpus(L, S):
    """BASELINE config 5 end points (2^19, 2^20 x 4096 B: golden hashes of `ntt n L 4096`) and two small orders vs the oracle."""
    import torch
    from fastecc_b200 import sharded
    world = min(torch.cuda.device_count(), 8)
    if world >= 2 and not sharded.p2p_ntt_supported(1 << L, 1 << (world.bit_length() - 1)):
        pytest.skip("this order cannot split its first-pass tiles over that many ranks")
    _run(_worker_p2p_ntt, L, S)


def test_p2p_fused_exchange_headline_hash_all_gpus():
    """N = 2^19 x 4096 B over every visible GPU (2, 4 or 8): golden parity hash 4272226309."""
    _run(_worker_p2p_headline, 19, 1024)


@pytest.mark.parametrize("L,S", [(11, 64), (16, 1024)])
def test_sharded_encode_on_gpus(L, S):
    _run(_worker, L, S)


def test_cpp_host_drives_the_sharded_encode(tmp_path):
    """integration/shard_example.cpp: one process per GPU, plain C++ on the C ABI (IPC handles exchanged through files, the
    passes and barriers through fastecc_b200_rs_encode_shard_p2p); every rank's parity rows must hash like the oracle's."""
    import subprocess
    import torch
    import oracle_lib as ol
    import fastecc_b200 as fe
    exe = os.path.join(ROOT, "integration", "shard_example")
    world = min(torch.cuda.device_count(), 8)
    if world < 2 or not os.path.exists(exe):
        pytest.skip("needs at least 2 GPUs and the built example (python __graft_entry__.py)")
    world = 1 << (world.bit_length() - 1)
    L, S = 16, 64
    procs = [subprocess.Popen([exe, str(r), str(world), str(L), str(S), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-500:] for o in outs]
    o = ol.load_oracle()
    want = ol.o_encode(o, ol.fill_A(o, 1 << L, S))
    for r in range(world):
        assert ("local hash %d" % fe.reference_hash(np.ascontiguousarray(want[r::world]))) in outs[r][0], outs[r][0]
