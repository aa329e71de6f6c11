"""fastecc_b200 -- B200 (sm_100a) NTT Reed-Solomon encoder behind FastECC's call surface.

Python host-side mirror of the reference interface for the one hot path:

    MFA_NTT(data, N, SIZE, InvNTT)          <->  template MFA_NTT<uint32_t,0xFFF00001>      ntt.cpp:382-383
    EncodeReedSolomon_body(data, N, SIZE)   <->  the timed body of EncodeReedSolomon        RS.cpp:41-63
    GF_Add / GF_Sub / GF_Mul / GF_Pow / GF_Root / GF_Inv  (host scalars, GF(p).cpp:37-48,110-127,254-297)

Everything that computes on blocks goes through the C ABI in include/fastecc_b200.h (ctypes); there is no CPU
fallback: if libfastecc_b200.so is missing, or no sm_100 GPU is usable, the calls raise.  `data` mirrors the
reference's ``T** data``: either a C-contiguous numpy uint32 array of shape (N, SIZE) (block i = data[i]) or a
sequence of N 1-D uint32 arrays (arbitrary block addresses).  Results are written in place.
"""
from __future__ import annotations

import ctypes
import os
from typing import Sequence, Union

import numpy as np

P = 0xFFF00001
MAX_LOG_N = 20
MAX_LOG_N_ENCODE = 19

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FASTECC_B200_LIB", os.path.join(_HERE, "libfastecc_b200.so"))   # override: A/B builds of the same ABI
_lib = None


class FastEccError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fastecc_b200 error {code}: {msg}")
        self.code = code


def lib() -> ctypes.CDLL:
    """Load the CUDA library (built by fastecc_b200.build / __graft_entry__.build). Never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -m fastecc_b200.build` (no CPU fallback exists)")
        L = ctypes.CDLL(LIB_PATH)
        sz, vp, ci = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int
        L.fastecc_b200_init.argtypes = [ci]; L.fastecc_b200_init.restype = ci
        L.fastecc_b200_shutdown.argtypes = []; L.fastecc_b200_shutdown.restype = None
        L.fastecc_b200_last_error.argtypes = []; L.fastecc_b200_last_error.restype = ctypes.c_char_p
        L.fastecc_b200_device.argtypes = []; L.fastecc_b200_device.restype = ci
        L.fastecc_b200_num_sms.argtypes = []; L.fastecc_b200_num_sms.restype = ci
        L.fastecc_b200_ntt_u32.argtypes = [vp, sz, sz, ci]; L.fastecc_b200_ntt_u32.restype = ci
        L.fastecc_b200_rs_encode.argtypes = [vp, sz, sz]; L.fastecc_b200_rs_encode.restype = ci
        L.fastecc_b200_ntt_u32_dev.argtypes = [vp, sz, sz, sz, ci, vp]; L.fastecc_b200_ntt_u32_dev.restype = ci
        L.fastecc_b200_rs_encode_dev.argtypes = [vp, sz, sz, sz, vp]; L.fastecc_b200_rs_encode_dev.restype = ci
        L.fastecc_b200_rs_encode_asym.argtypes = [vp, sz, sz, sz]; L.fastecc_b200_rs_encode_asym.restype = ci
        L.fastecc_b200_rs_encode_asym_dev.argtypes = [vp, sz, sz, sz, sz, vp]; L.fastecc_b200_rs_encode_asym_dev.restype = ci
        L.fastecc_b200_bytes_to_gfp_dev.argtypes = [vp, vp, sz, sz, sz, vp]; L.fastecc_b200_bytes_to_gfp_dev.restype = ci
        L.fastecc_b200_gfp_to_bytes_dev.argtypes = [vp, vp, sz, sz, sz, vp]; L.fastecc_b200_gfp_to_bytes_dev.restype = ci
        L.fastecc_b200_gf_mul_dev.argtypes = [vp, vp, vp, sz, vp]; L.fastecc_b200_gf_mul_dev.restype = ci
        L.fastecc_b200_gf_inv_dev.argtypes = [vp, vp, sz, vp]; L.fastecc_b200_gf_inv_dev.restype = ci
        L.fastecc_b200_row_scale_dev.argtypes = [vp, sz, sz, sz, vp, vp]; L.fastecc_b200_row_scale_dev.restype = ci
        L.fastecc_b200_rs_decode_pattern.argtypes = [sz, vp, sz, vp, ctypes.POINTER(vp)]; L.fastecc_b200_rs_decode_pattern.restype = ci
        L.fastecc_b200_rs_decode_recover.argtypes = [vp, vp, sz, sz, vp, sz, vp]; L.fastecc_b200_rs_decode_recover.restype = ci
        L.fastecc_b200_rs_decode_count.argtypes = [vp]; L.fastecc_b200_rs_decode_count.restype = sz
        L.fastecc_b200_rs_decode_free.argtypes = [vp]; L.fastecc_b200_rs_decode_free.restype = None
        L.fastecc_b200_rs_encode_shard_pass.argtypes = [vp, sz, ci, ci, sz, sz, ci, vp]; L.fastecc_b200_rs_encode_shard_pass.restype = ci
        L.fastecc_b200_rs_encode_shard_pass_p2p.argtypes = [vp, vp, sz, ci, ci, sz, sz, ci, vp]; L.fastecc_b200_rs_encode_shard_pass_p2p.restype = ci
        L.fastecc_b200_ntt_shard_pass_p2p.argtypes = [vp, vp, sz, ci, ci, sz, sz, ci, ci, vp]; L.fastecc_b200_ntt_shard_pass_p2p.restype = ci
        L.fastecc_b200_rs_encode_shard_p2p.argtypes = [vp, vp, vp, vp, sz, ci, ci, sz, sz, vp]; L.fastecc_b200_rs_encode_shard_p2p.restype = ci
        L.fastecc_b200_ntt_shard_p2p.argtypes = [vp, vp, vp, vp, sz, ci, ci, sz, sz, ci, vp]; L.fastecc_b200_ntt_shard_p2p.restype = ci
        L.fastecc_b200_shard_barrier.argtypes = [vp, ci, ci, ctypes.c_uint32, vp]; L.fastecc_b200_shard_barrier.restype = ci
        L.fastecc_b200_dev_alloc.argtypes = [sz]; L.fastecc_b200_dev_alloc.restype = vp
        L.fastecc_b200_dev_free.argtypes = [vp]; L.fastecc_b200_dev_free.restype = None
        L.fastecc_b200_ipc_export.argtypes = [vp, vp]; L.fastecc_b200_ipc_export.restype = ci
        L.fastecc_b200_ipc_open.argtypes = [vp, ctypes.POINTER(vp)]; L.fastecc_b200_ipc_open.restype = ci
        L.fastecc_b200_ipc_close.argtypes = [vp]; L.fastecc_b200_ipc_close.restype = ci
        L.fastecc_b200_kernel_launches.argtypes = []; L.fastecc_b200_kernel_launches.restype = ctypes.c_ulonglong
        L.fastecc_b200_rs_encode_dev_timed.argtypes = [vp, sz, sz, sz, vp, vp, vp, vp]; L.fastecc_b200_rs_encode_dev_timed.restype = ci
        L.fastecc_b200_shard_geometry.argtypes = [sz, ci, vp, vp, vp]; L.fastecc_b200_shard_geometry.restype = ci
        L.fastecc_b200_copy2d_async.argtypes = [vp, sz, vp, sz, sz, sz, ci, vp]; L.fastecc_b200_copy2d_async.restype = ci
        L.fastecc_b200_pin_host_buffers.argtypes = [ci]; L.fastecc_b200_pin_host_buffers.restype = ci
        L.fastecc_b200_hash_u32.argtypes = [vp, sz, sz]; L.fastecc_b200_hash_u32.restype = ctypes.c_uint32
        L.fastecc_b200_host_alloc.argtypes = [sz]; L.fastecc_b200_host_alloc.restype = vp
        L.fastecc_b200_host_free.argtypes = [vp]; L.fastecc_b200_host_free.restype = None
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise FastEccError(rc, lib().fastecc_b200_last_error().decode())


def init(device: int = 0) -> None:
    _check(lib().fastecc_b200_init(int(device)))


def shutdown() -> None:
    lib().fastecc_b200_shutdown()


def pin_host_buffers(enable: bool = True) -> None:
    """Opt in to page-locking large pageable arrays passed to MFA_NTT / EncodeReedSolomon_body in place (see the header)."""
    _check(lib().fastecc_b200_pin_host_buffers(1 if enable else 0))


def kernel_launches() -> int:
    return int(lib().fastecc_b200_kernel_launches())


# ---- host scalars of the reference call surface (pure Python ints; used by callers such as RS.cpp:51-57) ----------
def GF_Sub(x: int, y: int) -> int:            # GF(p).cpp:37-42
    return (x - y) % P


def GF_Add(x: int, y: int) -> int:            # GF(p).cpp:44-48
    return (x + y) % P


def GF_Mul(x: int, y: int) -> int:            # GF(p).cpp:110-127
    return (x * y) % P


def GF_Pow(x: int, n: int) -> int:            # GF(p).cpp:254-264
    return pow(x, n, P)


def GF_Root(n: int) -> int:                   # GF(p).cpp:267-276
    return pow(19, (P - 1) // n, P)


def GF_Inv(x: int) -> int:                    # GF(p).cpp:293-297
    return pow(x, P - 2, P)


Blocks = Union[np.ndarray, Sequence[np.ndarray]]


def _pointer_table(data: Blocks, N: int, SIZE: int):
    """Build the reference's ``T** data`` (RS.cpp:31-33) for a 2-D array or a sequence of block arrays."""
    if isinstance(data, np.ndarray):
        if data.dtype != np.uint32 or data.ndim != 2 or data.shape != (N, SIZE) or not data.flags.c_contiguous or not data.flags.writeable:
            raise ValueError("data must be a writeable C-contiguous uint32 array of shape (N, SIZE)")
        tab = data.ctypes.data + np.arange(N, dtype=np.uint64) * np.uint64(SIZE * 4)      # data[i] = data0 + i*SIZE
        return tab, data
    if len(data) != N:
        raise ValueError("len(data) != N")
    tab = np.empty(N, dtype=np.uint64)
    for i, blk in enumerate(data):
        if blk.dtype != np.uint32 or blk.ndim != 1 or blk.shape[0] < SIZE or not blk.flags.c_contiguous or not blk.flags.writeable:
            raise ValueError(f"block {i} must be a writeable contiguous uint32 vector of >= SIZE words")
        tab[i] = blk.ctypes.data
    return tab, data


def MFA_NTT(data: Blocks, N: int, SIZE: int, InvNTT: bool) -> None:
    """In-place length-N NTT of every word column (unnormalised inverse), bit-exact with ntt.cpp:382-447."""
    tab, keep = _pointer_table(data, N, SIZE)
    _check(lib().fastecc_b200_ntt_u32(tab.ctypes.data, N, SIZE, 1 if InvNTT else 0))
    del keep


def EncodeReedSolomon_body(data: Blocks, N: int, SIZE: int) -> None:
    """N data blocks -> N parity blocks in place: MFA_NTT(inv); x[i] *= root_2N^i/N; MFA_NTT(fwd)  (RS.cpp:41-63)."""
    tab, keep = _pointer_table(data, N, SIZE)
    _check(lib().fastecc_b200_rs_encode(tab.ctypes.data, N, SIZE))
    del keep


def EncodeReedSolomon_asym(data: Blocks, N: int, M: int, SIZE: int) -> None:
    """N data blocks -> M = N/2^k parity blocks in data[0..M): parity block j' = block (N/M)*j' of the full encode
    (the even points of the second transform, RS.cpp:65-66)."""
    tab, keep = _pointer_table(data, N, SIZE)
    _check(lib().fastecc_b200_rs_encode_asym(tab.ctypes.data, N, M, SIZE))
    del keep


# ---- device-resident entry points (torch tensors are used only as owners of device memory) -----------------------
def _dev_args(t):
    import torch
    if not (t.is_cuda and t.dim() == 2 and t.dtype in (torch.int32, torch.uint32) and t.stride(1) == 1):
        raise ValueError("expected a 2-D CUDA int32/uint32 tensor with unit stride along the word dimension")
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), torch.cuda.current_stream(t.device).cuda_stream


def ntt_dev(t, inverse: bool = False) -> None:
    ptr, N, size, pitch, stream = _dev_args(t)
    _check(lib().fastecc_b200_ntt_u32_dev(ptr, N, size, pitch, 1 if inverse else 0, stream))


def rs_encode_dev(t) -> None:
    ptr, N, size, pitch, stream = _dev_args(t)
    _check(lib().fastecc_b200_rs_encode_dev(ptr, N, size, pitch, stream))


def rs_encode_dev_timed(t):
    """One encode with an event between the passes: [(kernel instantiation, ms), ...] (synchronises the stream)."""
    ptr, N, size, pitch, stream = _dev_args(t)
    cap = 8
    ms = (ctypes.c_float * cap)(); names = (ctypes.c_char_p * cap)(); n = ctypes.c_int(cap)
    _check(lib().fastecc_b200_rs_encode_dev_timed(ptr, N, size, pitch, stream, ms, names, ctypes.byref(n)))
    return [((names[i] or b"?").decode(), float(ms[i])) for i in range(n.value)]


def rs_encode_asym_dev(t, M: int) -> None:
    """Rows [0, M) of t receive the M parity blocks; the other rows are undefined afterwards."""
    ptr, N, size, pitch, stream = _dev_args(t)
    _check(lib().fastecc_b200_rs_encode_asym_dev(ptr, N, M, size, pitch, stream))


def bytes_to_gfp_dev(raw, words=None):
    """raw: 2-D CUDA uint8 tensor [n_blocks, 4*W] (contiguous) -> int32 tensor [n_blocks, pitch] with words 0..W of every row
    set (GF.md:72-104 recoding: all of them < P).  pitch = W + 4 keeps rows 16-byte aligned."""
    import torch
    n, nbytes = raw.shape
    W = nbytes // 4
    if words is None:
        words = torch.zeros((n, W + 4), dtype=torch.int32, device=raw.device)
    _check(lib().fastecc_b200_bytes_to_gfp_dev(raw.data_ptr(), words.data_ptr(), n, W, words.stride(0), torch.cuda.current_stream(raw.device).cuda_stream))
    return words


def gfp_to_bytes_dev(words, W: int):
    import torch
    n = words.shape[0]
    raw = torch.empty((n, 4 * W), dtype=torch.uint8, device=words.device)
    _check(lib().fastecc_b200_gfp_to_bytes_dev(words.data_ptr(), raw.data_ptr(), n, W, words.stride(0), torch.cuda.current_stream(words.device).cuda_stream))
    return raw


def reference_hash(data: Blocks) -> int:
    """main.cpp:203-212 rolling hash over blocks in data[i] order (verification only; sequential by construction)."""
    if isinstance(data, np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint32)
        if data.ndim == 1:
            data = data.reshape(1, -1)
        N, SIZE = data.shape
    else:
        N, SIZE = len(data), min(int(b.shape[0]) for b in data)
    tab = (data.ctypes.data + np.arange(N, dtype=np.uint64) * np.uint64(SIZE * 4)) if isinstance(data, np.ndarray) \
        else np.array([b.ctypes.data for b in data], dtype=np.uint64)
    return int(lib().fastecc_b200_hash_u32(tab.ctypes.data, N, SIZE))
