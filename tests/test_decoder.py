"""Erasure decoder.  The orchestration in fastecc_b200/decoder.py (locator product tree, derivative
method) is the production code; on the CPU its five primitives are served by the oracle (tests only), on the GPU by the
C ABI.  Checked against the by-definition decoder (oracle/decode_oracle.py) and by encode -> erase -> decode == data."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import decode_oracle                                    # noqa: E402
import oracle_lib as ol                                 # noqa: E402

P = 0xFFF00001


class OracleBackend:
    """CPU stand-in for decoder.CudaBackend: torch CPU tensors in, oracle / numpy arithmetic."""

    def __init__(self, oracle):
        self.o = oracle

    @staticmethod
    def _u(t):
        return t.contiguous().numpy().view(np.uint32)

    def ntt(self, t, inverse):
        import torch
        a = np.ascontiguousarray(self._u(t))
        t.copy_(torch.from_numpy(ol.o_ntt(self.o, a, inverse).view(np.int32)))

    def gf_mul(self, a, b):
        import torch
        r = (self._u(a).astype(np.uint64) % P) * (self._u(b).astype(np.uint64) % P) % P
        return torch.from_numpy(r.astype(np.uint32).view(np.int32).reshape(tuple(a.shape)))

    def gf_inv(self, a):
        import torch
        r = np.array([pow(int(x) % P, P - 2, P) for x in self._u(a).ravel()], dtype=np.uint32)
        return torch.from_numpy(r.view(np.int32).reshape(tuple(a.shape)))

    def row_scale(self, t, consts):
        t.copy_(self.gf_mul(t, consts.view(-1, 1).expand_as(t).contiguous()))


def _codeword(oracle, N, S, seed):
    rng = np.random.default_rng(seed)
    data = rng.integers(0, P, size=(N, S), dtype=np.uint64).astype(np.uint32)
    parity = ol.o_encode(oracle, data.copy())
    code = np.empty((2 * N, S), dtype=np.uint32)
    code[0::2] = data
    code[1::2] = parity
    return code, rng


def _run(be, oracle, N, S, n_erased, seed, device="cpu"):
    import torch
    from fastecc_b200 import decoder
    code, rng = _codeword(oracle, N, S, seed)
    erased = sorted(rng.choice(2 * N, size=n_erased, replace=False).tolist())
    damaged = code.copy()
    damaged[erased] = 0xDEADBEEF
    t = torch.from_numpy(damaged.view(np.int32)).to(device)
    rec = decoder.decode(t, erased, be).cpu().numpy().view(np.uint32)
    assert np.array_equal(rec, code[erased])
    return code, erased, rec


@pytest.mark.parametrize("N,S,n_erased", [(1, 2, 1), (2, 3, 1), (4, 2, 4), (8, 4, 5), (16, 2, 16), (32, 3, 20), (64, 2, 64)])
def test_decoder_orchestration_on_oracle_backend(oracle, N, S, n_erased):
    code, erased, rec = _run(OracleBackend(oracle), oracle, N, S, n_erased, 7 * N + n_erased)
    if N <= 16:                                                # and the by-definition decoder agrees
        rows = [[int(v) for v in r] for r in code]
        want = decode_oracle.recover(rows, erased)
        for k, e in enumerate(erased):
            assert [int(v) for v in rec[k]] == want[e]


def test_locator_vanishes_exactly_on_the_erased_points(oracle):
    import torch
    from fastecc_b200 import decoder
    be = OracleBackend(oracle)
    n2, erased = 64, [0, 3, 4, 17, 31, 32, 63]
    lc = decoder.locator_coefficients(erased, n2, "cpu", be)
    lv = lc.clone().view(n2, 1)
    be.ntt(lv, False)
    zeros = set(np.flatnonzero(lv.numpy().view(np.uint32).ravel() == 0).tolist())
    assert zeros == set(erased)
    assert int(lc.numpy().view(np.uint32)[len(erased)]) == 1 and not lc.numpy()[len(erased) + 1:].any()     # monic, degree |E|


def test_decoder_argument_validation(oracle):
    import torch
    from fastecc_b200 import decoder
    be = OracleBackend(oracle)
    t = torch.zeros((16, 2), dtype=torch.int32)
    for bad in ([3, 3], [5, 2], [16], list(range(9))):
        with pytest.raises(ValueError):
            decoder.decode(t, bad, be)
    with pytest.raises(ValueError):
        decoder.decode(torch.zeros((12, 2), dtype=torch.int32), [1], be)


@pytest.mark.gpu
@pytest.mark.parametrize("N,S,n_erased", [(1, 4, 1), (4, 8, 3), (16, 4, 16), (64, 16, 40), (1024, 64, 1024), (4096, 32, 3000), (1 << 15, 16, 1 << 15)])
def test_gpu_decoder_recovers_erased_blocks(fecc, oracle, N, S, n_erased):
    from fastecc_b200 import decoder
    _run(decoder.CudaBackend(), oracle, N, S, n_erased, 1000 + N, device="cuda")


@pytest.mark.gpu
def test_gpu_full_size_encode_erase_decode(fecc, oracle):
    """BASELINE headline order: 2^19 data + 2^19 parity blocks (1 KiB blocks to bound memory), half of the code word
    erased at random, every lost data block must come back -- encoder and decoder both on the CUDA path."""
    import torch
    from fastecc_b200 import decoder
    N, S = 1 << 19, 256
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    data = torch.randint(0, P, (N, S), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    par = data.clone()
    fecc.rs_encode_dev(par)
    code = torch.empty((2 * N, S), dtype=torch.int32, device="cuda")
    code[0::2] = data
    code[1::2] = par
    del par
    perm = torch.randperm(2 * N, device="cuda", generator=g)
    erased = torch.sort(perm[:N]).values
    want = code[erased].clone()
    code[erased] = -1
    rec = decoder.decode(code, erased.cpu().tolist(), decoder.CudaBackend())
    assert bool((rec == want).all())


@pytest.mark.gpu
@pytest.mark.parametrize("N,S,n_erased", [(1, 4, 1), (4, 8, 3), (16, 4, 16), (32, 8, 5), (64, 16, 40), (1024, 64, 1024), (4096, 32, 3000), (1 << 15, 16, 1 << 15)])
def test_gpu_c_abi_decoder_recovers_erased_blocks(fecc, oracle, N, S, n_erased):
    """fastecc_b200_rs_decode_pattern / _recover (csrc/decode.cu): the decoder a C++ host calls."""
    import torch
    from fastecc_b200 import decoder
    code, rng = _codeword(oracle, N, S, 2000 + N)
    erased = sorted(rng.choice(2 * N, size=n_erased, replace=False).tolist())
    damaged = code.copy()
    damaged[erased] = 0xDEADBEEF
    pat = decoder.CErasurePattern(2 * N, erased)
    for _ in range(2):                                           # a pattern serves any number of code words
        rec = pat.recover(torch.from_numpy(damaged.view(np.int32)).cuda()).cpu().numpy().view(np.uint32)
        assert np.array_equal(rec, code[erased])
    pat.close()
    if N <= 16:
        want = decode_oracle.recover([[int(v) for v in r] for r in code], erased)
        for k, e in enumerate(erased):
            assert [int(v) for v in rec[k]] == want[e]


@pytest.mark.gpu
def test_gpu_c_abi_decoder_rejects_bad_patterns(fecc):
    from fastecc_b200 import decoder
    for n2, bad in ((16, [3, 3]), (16, [5, 2]), (16, [16]), (16, list(range(9))), (12, [1])):
        with pytest.raises(fecc.FastEccError):
            decoder.CErasurePattern(n2, bad)


@pytest.mark.gpu
def test_gpu_c_abi_full_size_encode_erase_decode(fecc, oracle):
    """2^19 data + 2^19 parity blocks of 1 KiB, half of the code word erased at random: encoder and C-ABI decoder."""
    import time
    import torch
    from fastecc_b200 import decoder
    N, S = 1 << 19, 256
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    data = torch.randint(0, P, (N, S), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    par = data.clone()
    fecc.rs_encode_dev(par)
    code = torch.empty((2 * N, S), dtype=torch.int32, device="cuda")
    code[0::2] = data
    code[1::2] = par
    del par
    perm = torch.randperm(2 * N, device="cuda", generator=g)
    erased = torch.sort(perm[:N]).values
    want = code[erased].clone()
    code[erased] = -1
    pos = erased.cpu().tolist()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pat = decoder.CErasurePattern(2 * N, pos)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec = pat.recover(code)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("C-ABI decoder, 2^19 of 2^20 rows erased: pattern %.1f ms (first use, incl. table builds), recover %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    assert bool((rec == want).all())
    pat.close()
