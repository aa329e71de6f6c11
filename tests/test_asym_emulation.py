"""CPU: asymmetric encode (N data -> M = N/2^k parity blocks).  The pass descriptors of
csrc/plan.h plan_encode_asym run on the CPU emulation of the kernel; the result must be every (N/M)-th parity block of
the oracle's full encode (the reference describes exactly this subset: RS.cpp:65-66)."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol                                   # noqa: E402
from test_sharded_gloo import EMU, _build_emulator        # noqa: E402


@pytest.mark.parametrize("L,K,S", [(11, 1, 8), (12, 2, 4), (13, 1, 12), (16, 3, 4), (19, 1, 4), (19, 10, 4)])
def test_asym_passes_match_full_encode_subset(L, K, S):
    _build_emulator()
    emu = ctypes.CDLL(EMU)
    f = emu.emu_rs_encode_asym
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
    N, M = 1 << L, 1 << (L - K)
    o = ol.load_oracle()
    full = ol.fill_B(o, N, S)
    want = ol.o_encode(o, full)[::N // M]
    x = full.copy()
    y = np.zeros_like(x)
    assert f(x.ctypes.data, y.ctypes.data, N, M, S, S) == 0
    assert np.array_equal(x[:M], want)


def test_asym_native_limits():
    _build_emulator()
    emu = ctypes.CDLL(EMU)
    f = emu.emu_rs_encode_asym
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
    x = np.zeros((1 << 12, 4), dtype=np.uint32)
    assert f(x.ctypes.data, x.ctypes.data, 1 << 12, 1 << 12, 4, 4) == -1    # M == N is the plain encode
    assert f(x.ctypes.data, x.ctypes.data, 1 << 12, 1 << 4, 4, 4) == -1     # M < N1: served by a gather in api.cu
    assert f(x.ctypes.data, x.ctypes.data, 1 << 10, 1 << 9, 4, 4) == -1     # single-pass orders: full encode + gather
