"""Byte <-> GF(p) recoding (GF.md:72-104 -- described, not implemented, by the reference).
CPU: properties of the restatement in oracle/byte_recode_oracle.py.  GPU: the CUDA kernels against it, bit for bit, and
the whole chain bytes -> words < P -> encode (1025-word blocks) -> parity -> ... -> bytes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import byte_recode_oracle as bro                    # noqa: E402

P = 0xFFF00001


def _cases(W, rng):
    blocks = [rng.integers(0, 1 << 32, size=W, dtype=np.uint64).astype(np.uint32) for _ in range(6)]
    clean = rng.integers(0, 0xFFF00000, size=W, dtype=np.uint64).astype(np.uint32)                     # no 0xFFF digit
    allf = np.full(W, 0xFFFFFFFF, dtype=np.uint32)                                                     # every digit is 0xFFF
    first = clean.copy(); first[0] = 0xFFF12345
    last = clean.copy(); last[W - 1] = 0xFFFFFFFF
    pm1 = clean.copy(); pm1[W // 2] = P - 1                                                            # a legal word whose digit is 0xFFF
    many = clean.copy(); many[::3] |= 0xFFF00000
    flagish = ((rng.integers(0x400, 0x800, size=W, dtype=np.uint64) << 20) | 5).astype(np.uint32)      # digits that look like "more follow" entries
    flagish[min(7, W - 1)] = 0xFFF00001
    return np.stack(blocks + [clean, allf, first, last, pm1, many, flagish])


@pytest.mark.parametrize("W", [4, 64, 512, 1024])
def test_oracle_recoding_properties(W):
    rng = np.random.default_rng(W)
    x = _cases(W, rng)
    enc = bro.bytes_to_gfp(x)
    assert enc.shape == (x.shape[0], W + 1)
    assert int(enc.max()) < P                                                 # everything encodable
    assert np.array_equal(bro.gfp_to_bytes(enc), x)                           # lossless
    clean = (x >> 20 != 0xFFF).all(axis=1)
    assert np.array_equal(enc[clean, :W], x[clean]) and not enc[clean, W].any()      # untouched when no digit is 0xFFF
    assert enc[~clean, W].all()
    assert np.array_equal(enc[:, :W] & 0xFFFFF, x & 0xFFFFF)                  # low 20 bits stay in place


@pytest.mark.gpu
@pytest.mark.parametrize("W", [4, 64, 516, 1024])
def test_gpu_recoding_matches_oracle(fecc, W):
    import torch
    rng = np.random.default_rng(50 + W)
    x = np.concatenate([_cases(W, rng), rng.integers(0, 1 << 32, size=(300, W), dtype=np.uint64).astype(np.uint32)])
    raw = torch.from_numpy(x.view(np.uint8).reshape(x.shape[0], 4 * W)).cuda()
    words = fecc.bytes_to_gfp_dev(raw)
    got = words.cpu().numpy().view(np.uint32)
    want = bro.bytes_to_gfp(x)
    assert np.array_equal(got[:, :W + 1], want)
    back = fecc.gfp_to_bytes_dev(words, W)
    assert bool((back == raw).all())


@pytest.mark.gpu
def test_gpu_bytes_encode_roundtrip(fecc, oracle):
    """Arbitrary bytes -> recoded 1025-word blocks -> RS encode through the C ABI == oracle encode of the oracle's recoding;
    data recovered from the recoded blocks byte for byte."""
    import torch
    import oracle_lib as ol
    N, W = 256, 1024
    rng = np.random.default_rng(9)
    x = rng.integers(0, 1 << 32, size=(N, W), dtype=np.uint64).astype(np.uint32)
    raw = torch.from_numpy(x.view(np.uint8).reshape(N, 4 * W)).cuda()
    words = fecc.bytes_to_gfp_dev(raw)                                         # [N, 1028]
    keep = words.clone()
    fecc.rs_encode_dev(words[:, :W + 1])                                       # SIZE = 1025 words, pitch 1028
    want = ol.o_encode(oracle, np.ascontiguousarray(bro.bytes_to_gfp(x)))
    assert np.array_equal(words.cpu().numpy().view(np.uint32)[:, :W + 1], want)
    assert bool((fecc.gfp_to_bytes_dev(keep, W) == raw).all())


@pytest.mark.gpu
def test_gpu_recoding_argument_validation(fecc):
    import torch
    raw = torch.zeros((4, 4096), dtype=torch.uint8, device="cuda")
    bad = torch.zeros((4, 1024), dtype=torch.int32, device="cuda")             # no room for the extra word
    with pytest.raises(Exception):
        fecc.bytes_to_gfp_dev(raw, bad)
    with pytest.raises(Exception):
        fecc.bytes_to_gfp_dev(torch.zeros((4, 4 * 1028), dtype=torch.uint8, device="cuda"))   # W > 1024
