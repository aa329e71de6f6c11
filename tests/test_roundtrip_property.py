"""Encode -> erase -> decode round trip: the systematic codeword of the encoder is c[2i] = data[i], c[2j+1] = parity[j]
with c[m] = f(rho^m), rho = GF_Root(2N), deg f < N (RS.cpp:22-68).  Any N of the 2N symbols determine
f, so erasing N random symbols and interpolating the rest (plain Lagrange in Python integers -- the reference has no
decoder) must give back every erased data block.  Run against the oracle on the CPU and against the CUDA path on a GPU."""
import numpy as np
import pytest

import oracle_lib as ol

P = 0xFFF00001


def _recover(code, keep, targets, rho):
    """code: dict position -> vector of words (python ints); Lagrange-evaluate f at rho^t for t in targets."""
    xs = [pow(rho, m, P) for m in keep]
    denom_inv = []
    for a, xa in enumerate(xs):
        d = 1
        for b, xb in enumerate(xs):
            if a != b:
                d = d * (xa - xb) % P
        denom_inv.append(pow(d, P - 2, P))
    out = {}
    for t in targets:
        xt = pow(rho, t, P)
        diffs = [(xt - xb) % P for xb in xs]
        acc = [0] * len(code[keep[0]])
        for a in range(len(xs)):
            num = 1
            for b, df in enumerate(diffs):
                if a != b:
                    num = num * df % P
            coef = num * denom_inv[a] % P
            va = code[keep[a]]
            acc = [(s + coef * v) % P for s, v in zip(acc, va)]
        out[t] = acc
    return out


def _check(encode, oracle, L, S, seed):
    N = 1 << L
    rng = np.random.default_rng(seed)
    data = rng.integers(0, P, size=(N, S), dtype=np.uint64).astype(np.uint32)
    parity = encode(data.copy())
    code = {}
    for i in range(N):
        code[2 * i] = [int(v) for v in data[i]]
        code[2 * i + 1] = [int(v) for v in parity[i]]
    rho = pow(19, (P - 1) // (2 * N), P)                                   # GF_Root(2N), GF(p).cpp:267-276
    erased = sorted(rng.choice(2 * N, size=N, replace=False).tolist())
    keep = [m for m in range(2 * N) if m not in set(erased)]
    lost_data = [m for m in erased if m % 2 == 0]
    rec = _recover(code, keep, lost_data, rho)
    for m in lost_data:
        assert rec[m] == code[m], "data block %d not recovered" % (m // 2)
    assert len(lost_data) > 0


@pytest.mark.parametrize("L,S", [(1, 3), (3, 4), (5, 4), (6, 2)])
def test_oracle_encode_erase_decode(oracle, L, S):
    _check(lambda a: ol.o_encode(oracle, a), oracle, L, S, 100 + L)


@pytest.mark.gpu
@pytest.mark.parametrize("L,S", [(2, 4), (4, 8), (5, 4), (6, 4), (7, 1)])
def test_gpu_encode_erase_decode(fecc, oracle, L, S):
    import torch

    def enc(a):
        t = torch.from_numpy(a.view(np.int32)).cuda()
        fecc.rs_encode_dev(t)
        return t.cpu().numpy().view(np.uint32)
    _check(enc, oracle, L, S, 200 + L)
