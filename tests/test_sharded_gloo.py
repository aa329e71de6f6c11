"""CPU, world_size 2 (and 4), gloo: the sharded single-transform encoder (BASELINE config 4).  The orchestration
(fastecc_b200/sharded.py: cyclic block ownership, two all-to-alls, local pass order) and the sharded pass descriptors
(csrc/plan.h plan_encode_shard) are the production ones; only the local pass runs on the CPU emulation of the
kernel instead of the GPU.  Rank 0 reassembles the parity and compares it with the oracle bit for bit."""
import ctypes
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
EMU = os.path.join(HERE, "libemulate.so")


def _build_emulator():
    src = [os.path.join(HERE, "emulate_lib.cu"), os.path.join(HERE, "emulate_pass.h"), os.path.join(ROOT, "fastecc_b200", "csrc", "ntt_tile.cuh"),
           os.path.join(ROOT, "fastecc_b200", "csrc", "plan.h")]
    if os.path.exists(EMU) and all(os.path.getmtime(EMU) > os.path.getmtime(f) for f in src):
        return
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-Wno-deprecated-gpu-targets", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", EMU, src[0],
                    "-I" + os.path.join(ROOT, "fastecc_b200", "csrc")], check=True, capture_output=True)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, L, S, q):
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    from fastecc_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = ctypes.CDLL(EMU)
    emu.emu_rs_encode_shard_pass.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
    N = 1 << L
    o = ol.load_oracle()
    full = ol.fill_B(o, N, S)
    local = torch.from_numpy(np.ascontiguousarray(full[rank::world]).view(np.int32))      # cyclic ownership

    def run_pass(t, which):
        assert t.is_contiguous()
        assert emu.emu_rs_encode_shard_pass(t.data_ptr(), N, world, rank, S, S, which) == 0

    out = sharded.rs_encode_sharded(local, N, world, run_pass)
    gathered = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, gathered, dst=0)
    if rank == 0:
        par = np.empty((N, S), dtype=np.uint32)
        for r in range(world):
            par[r::world] = gathered[r].numpy().view(np.uint32)
        q.put(bool(np.array_equal(par, ol.o_encode(o, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,L,S", [(2, 11, 8), (4, 12, 4)])
def test_sharded_encode_matches_oracle(world, L, S):
    import torch.multiprocessing as mp
    _build_emulator()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
