"""One Reed-Solomon encode sharded over G GPUs (BASELINE config 4).

Two implementations of the same decomposition:
  * P2PShardedEncoder -- the exchange is FUSED into the pass kernels: passes A and BC store every output row straight
    into the memory of the GPU that owns it (CUDA IPC mappings, NVLink), so there is no all-to-all, no pack/unpack
    and no extra trip through HBM; the passes are separated by a flag barrier over the same peer mappings (a one-warp
    kernel on the stream, fastecc_b200_shard_barrier).  torch.distributed only carries the IPC handles at set-up.
  * rs_encode_sharded -- local passes + two NCCL all-to-alls of whole blocks (the plain-library baseline; also what the
    CPU/gloo tests run, with the kernel emulated).

Blocks are dealt cyclically (global block i = l*G + rank is local block l) for the data going in and the parity
coming out.  N = N1*N2 as in csrc/plan.h; with N2 % G == 0 the row sets of pass A (fixed n2) and pass D (fixed j2)
are local to rank n2 % G / j2 % G, and those of the fused pass BC (fixed k1) to rank k1 % G, so the transpose
between them (TransposeMatrix on block pointers in the reference, ntt.cpp:322-341,415,433,445) becomes an all-to-all
of 4 KiB blocks over NVLink.  `run_pass(tensor, which)` executes one local pass in place: on the GPUs
fastecc_b200_rs_encode_shard_pass, in the CPU tests the emulation of the same kernel."""
from __future__ import annotations

from typing import Callable


def _geometry(N: int, G: int):
    """(N1, N2, fused_exchange_ok) from the planner itself (csrc/plan.h split_l1 / shard_supported through the C ABI), so
    that the exchanges below can never disagree with the passes' decomposition (e.g. under FASTECC_B200_SPLIT)."""
    import ctypes
    import fastecc_b200 as fe
    LN = N.bit_length() - 1
    if N != 1 << LN or LN < 11 or LN > 19 or G < 2 or G & (G - 1):
        raise ValueError("sharded encode needs N = 2^11..2^19 and a power-of-two number of ranks")
    n1, n2, ok = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_int()
    if fe.lib().fastecc_b200_shard_geometry(N, G, ctypes.byref(n1), ctypes.byref(n2), ctypes.byref(ok)) != 0:
        raise ValueError("N2 = %d must be a multiple of the number of ranks" % n2.value)
    return n1.value, n2.value, bool(ok.value & 1)


def geometry(N: int, G: int):
    return _geometry(N, G)[:2]


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


def _support_bits(N: int, G: int) -> int:
    """bit 0: the fused-exchange ENCODE passes, bit 1: the fused-exchange NTT passes support (N, G)  (csrc/plan.h)."""
    import ctypes
    import fastecc_b200 as fe
    ok = ctypes.c_int()
    if N < 2 or N & (N - 1) or G < 2 or G & (G - 1):
        return 0
    fe.lib().fastecc_b200_shard_geometry(N, G, None, None, ctypes.byref(ok))
    return ok.value


def p2p_supported(N: int, G: int) -> bool:
    """csrc/plan.h shard_p2p_supported(): <= 8 ranks and every thread's 32 output rows on one rank for both tile heights."""
    return bool(_support_bits(N, G) & 1)


def p2p_ntt_supported(N: int, G: int) -> bool:
    """csrc/plan.h ntt_shard_p2p_supported(): N = 2^11..2^20, <= 8 ranks, first tile height >= 32 * ranks."""
    return bool(_support_bits(N, G) & 2)


BARRIER_WORDS = 16                    # include/fastecc_b200.h FASTECC_B200_BARRIER_WORDS


class _RawCudaBuffer:
    """A cudaMalloc'ed block exposed through __cuda_array_interface__ (torch.as_tensor keeps this object alive)."""
    def __init__(self, ptr: int, shape, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class P2PShardedEncoder:
    """ONE encode of N blocks of S words over the G ranks of `group`, exchange fused into the kernels' stores.

    enc.x is this rank's [N/G, S] int32 CUDA tensor: fill it with the local data blocks (global block l*G + rank is row l),
    call enc.encode(), read the local parity blocks from the same tensor.  Collective: every rank constructs it and calls
    encode() the same number of times.  One process per GPU on one node (CUDA IPC)."""

    def __init__(self, N: int, S: int, group=None, barrier: str = "flags"):
        """barrier: "flags" (default: fastecc_b200_shard_barrier, a kernel over the peer mappings) or "nccl" (a one-word all-reduce)."""
        import ctypes
        import torch
        import torch.distributed as dist
        import fastecc_b200 as fe
        self.N, self.S, self.group = N, S, group
        if barrier not in ("flags", "nccl"):
            raise ValueError("barrier must be 'flags' or 'nccl'")
        self.barrier_kind = barrier
        self.G, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._can_encode, self._can_ntt = p2p_supported(N, self.G), p2p_ntt_supported(N, self.G)
        if not (self._can_encode or self._can_ntt) or S % 4:
            raise ValueError("fused-exchange sharding needs <= 8 ranks, SIZE % 4 == 0 and N = 2^15..2^19 (encode; fewer ranks: from 2^12) or up to 2^20 (NTT)")
        self._lib = L = fe.lib()
        self._dev = torch.cuda.current_device()
        fe.init(self._dev)
        rows = N // self.G
        nbytes = rows * S * 4
        self._own = [L.fastecc_b200_dev_alloc(nbytes), L.fastecc_b200_dev_alloc(nbytes), L.fastecc_b200_dev_alloc(4 * BARRIER_WORDS)]     # X, Y, barrier flags
        if not all(self._own):
            raise MemoryError(L.fastecc_b200_last_error().decode())
        self._flags = torch.as_tensor(_RawCudaBuffer(self._own[2], (BARRIER_WORDS,)), device=torch.device("cuda", self._dev))
        self._flags.zero_()
        torch.cuda.synchronize()                                       # the flags are zero before anybody can reach them
        self._epoch = ctypes.c_uint32(0)                               # barrier counter, shared with the single-call C entry points
        nh = len(self._own)
        handles = torch.empty(64 * nh, dtype=torch.uint8)
        hb = (ctypes.c_ubyte * (64 * nh))()
        for k in range(nh):
            fe._check(L.fastecc_b200_ipc_export(self._own[k], ctypes.addressof(hb) + 64 * k))
        handles.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
        dev_handles = handles.cuda()
        everyone = torch.empty(64 * nh * self.G, dtype=torch.uint8, device=dev_handles.device)
        dist.all_gather_into_tensor(everyone, dev_handles, group=group)
        everyone = everyone.cpu().numpy().reshape(self.G, nh, 64)
        self._opened = []
        peers = [[0] * self.G for _ in range(nh)]
        for r in range(self.G):
            for k in range(nh):
                if r == self.rank:
                    peers[k][r] = self._own[k]
                else:
                    p = ctypes.c_void_p()
                    fe._check(L.fastecc_b200_ipc_open(everyone[r, k].ctypes.data, ctypes.byref(p)))
                    peers[k][r] = p.value
                    self._opened.append(p.value)
        self._peers = peers                                            # [X, Y or flags][rank] -> device address in this process
        self._flag_peers = (ctypes.c_void_p * self.G)(*peers[2])
        self._ptr_cache = {}
        self.x = torch.as_tensor(_RawCudaBuffer(self._own[0], (rows, S)), device=torch.device("cuda", self._dev))
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.x.device)
        self._s_h2d = self._s_d2h = None
        self._nccl_barrier()                                           # nobody stores into a peer before everybody has mapped everything

    def _nccl_barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)                  # set-up / tear-down only

    def _barrier(self):
        """Orders the passes of all ranks on the current stream: a one-warp kernel that raises this rank's flag in every peer and
        waits for theirs (fastecc_b200_shard_barrier) -- no NCCL and no host involvement on the data path."""
        import torch
        import fastecc_b200 as fe
        if self.barrier_kind == "nccl":
            return self._nccl_barrier()
        self._epoch.value = (self._epoch.value + 1) & 0xFFFFFFFF
        fe._check(self._lib.fastecc_b200_shard_barrier(self._flag_peers, self.G, self.rank, self._epoch.value, torch.cuda.current_stream().cuda_stream))

    def check(self):
        """Raise if a barrier ever timed out (a peer died or fell out of step).  Synchronises the device."""
        import torch
        torch.cuda.synchronize()
        if int(self._flags[8].item()) != 0:
            raise RuntimeError("fastecc_b200 sharded barrier timed out on rank %d: the ranks are out of step" % self.rank)

    def _peer_arrays(self, off_bytes: int):
        """HOST arrays of the peers' X and Y addresses, shifted to a column offset (column chunks are independent codewords)."""
        import ctypes
        if off_bytes not in self._ptr_cache:
            self._ptr_cache[off_bytes] = tuple((ctypes.c_void_p * self.G)(*[p + off_bytes for p in self._peers[k]]) for k in range(2))
        return self._ptr_cache[off_bytes]

    def _passes(self, col0: int, width: int, events=None):
        """The three passes on word columns [col0, col0 + width) of every block, on the current stream."""
        import torch
        import fastecc_b200 as fe
        L, st = self._lib, torch.cuda.current_stream().cuda_stream
        X, Y = self._own[0] + 4 * col0, self._own[1] + 4 * col0
        xp, yp = self._peer_arrays(4 * col0)

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); events.append(e)
        mark()
        fe._check(L.fastecc_b200_rs_encode_shard_pass_p2p(X, yp, self.N, self.G, self.rank, width, self.S, 0, st))
        mark()
        self._barrier()
        mark()
        fe._check(L.fastecc_b200_rs_encode_shard_pass_p2p(Y, xp, self.N, self.G, self.rank, width, self.S, 1, st))
        mark()
        self._barrier()
        mark()
        fe._check(L.fastecc_b200_rs_encode_shard_pass_p2p(X, xp, self.N, self.G, self.rank, width, self.S, 2, st))
        mark()

    def encode(self, events=None):
        """events: optional list that receives 6 CUDA events bracketing pass A, barrier, pass BC, barrier, pass D."""
        import ctypes
        import torch
        import fastecc_b200 as fe
        if not self._can_encode:
            raise ValueError("N = %d cannot be encoded over %d ranks with the fused exchange" % (self.N, self.G))
        if events is None and self.barrier_kind == "flags":            # the whole rank-local sequence as ONE C call
            xp, yp = self._peer_arrays(0)
            fe._check(self._lib.fastecc_b200_rs_encode_shard_p2p(xp, yp, self._flag_peers, ctypes.byref(self._epoch), self.N, self.G, self.rank,
                                                                self.S, self.S, torch.cuda.current_stream().cuda_stream))
        else:
            self._passes(0, self.S, events)
        return self.x

    def ntt(self, inverse: bool = False, events=None):
        """ONE standalone transform of the N blocks (MFA_NTT, ntt.cpp:382-447; unnormalised inverse), in place in enc.x, cyclic
        blocks in and out: pass A' stores into the owners' Y over NVLink, barrier, pass B' is local, barrier (nobody may overwrite a Y
        its owner is still reading; the encode's passes have that barrier between BC and D already).  events: 4 CUDA events."""
        import torch
        import fastecc_b200 as fe
        if not self._can_ntt:
            raise ValueError("N = %d cannot be transformed over %d ranks with the fused exchange" % (self.N, self.G))
        import ctypes
        L, st = self._lib, torch.cuda.current_stream().cuda_stream
        X, Y = self._own[:2]
        xp, yp = self._peer_arrays(0)
        if events is None and self.barrier_kind == "flags":
            fe._check(L.fastecc_b200_ntt_shard_p2p(xp, yp, self._flag_peers, ctypes.byref(self._epoch), self.N, self.G, self.rank, self.S, self.S,
                                                   1 if inverse else 0, st))
            return self.x

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); events.append(e)
        mark()
        fe._check(L.fastecc_b200_ntt_shard_pass_p2p(X, yp, self.N, self.G, self.rank, self.S, self.S, 1 if inverse else 0, 0, st))
        mark()
        self._barrier()
        mark()
        fe._check(L.fastecc_b200_ntt_shard_pass_p2p(Y, xp, self.N, self.G, self.rank, self.S, self.S, 1 if inverse else 0, 1, st))
        mark()
        self._barrier()                        # the next operation's first pass stores into the peers' Y: they must be done reading it
        return self.x

    def encode_host(self, h_in, h_out=None, chunk_words: int = 256):
        """End to end from host memory: h_in is this rank's [N/G, S] uint32 numpy array (global block l*G + rank = row l),
        ideally pinned (fastecc_b200_host_alloc); the parity rows are written to h_out (default: in place).  The word
        columns are independent codewords, so the array moves in column chunks: H2D of chunk c+1, the three sharded passes
        of chunk c and D2H of chunk c-1 overlap on three streams.  Collective, blocking."""
        import torch
        import fastecc_b200 as fe
        rows, S = self.N // self.G, self.S
        h_out = h_in if h_out is None else h_out
        for h in (h_in, h_out):
            if h.dtype.itemsize != 4 or h.shape != (rows, S) or not h.flags.c_contiguous:
                raise ValueError("host shard must be a C-contiguous [N/G, S] array of 32-bit words")
        if chunk_words % 4 or chunk_words <= 0:
            raise ValueError("chunk_words must be a positive multiple of 4")
        if self._s_h2d is None:
            self._s_h2d, self._s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        L = self._lib
        src, dst, X = h_in.ctypes.data, h_out.ctypes.data, self._own[0]
        self._s_h2d.wait_stream(cur)                                    # earlier work on the buffers is done before we overwrite them
        for c0 in range(0, S, chunk_words):
            w = min(chunk_words, S - c0)
            fe._check(L.fastecc_b200_copy2d_async(X + 4 * c0, 4 * S, src + 4 * c0, 4 * S, 4 * w, rows, 1, self._s_h2d.cuda_stream))
            ev = torch.cuda.Event(); ev.record(self._s_h2d); cur.wait_event(ev)
            self._passes(c0, w)
            ev = torch.cuda.Event(); ev.record(cur); self._s_d2h.wait_event(ev)
            fe._check(L.fastecc_b200_copy2d_async(dst + 4 * c0, 4 * S, X + 4 * c0, 4 * S, 4 * w, rows, 0, self._s_d2h.cuda_stream))
        self._s_d2h.synchronize()
        cur.synchronize()
        return h_out

    def close(self):
        import torch
        if self._own:
            torch.cuda.synchronize()
            self._nccl_barrier(); torch.cuda.synchronize()             # no peer is still storing into our buffers
            timed_out = int(self._flags[8].item()) != 0
            for p in self._opened:
                self._lib.fastecc_b200_ipc_close(p)
            self.x = self._flags = None
            for p in self._own:
                self._lib.fastecc_b200_dev_free(p)
            self._own, self._opened = [], []
            if timed_out:
                raise RuntimeError("fastecc_b200 sharded barrier timed out on rank %d: the ranks were out of step, results are invalid" % self.rank)
