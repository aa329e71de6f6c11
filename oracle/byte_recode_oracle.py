"""TEST INFRASTRUCTURE ONLY (see oracle/gfp_oracle.h): CPU restatement of the byte <-> GF(p) recoding that the reference
DESCRIBES but does not implement (GF.md:72-104 "Efficient data packing", README.md:160-163).  Parity is therefore
UNPINNED by the reference: there is no golden vector and no reference code to run; the restatement follows the prose
literally and is checked by its own properties (round trip, range, identity on blocks without a 0xFFF digit).

Block = W <= 1024 32-bit words; digit(i) = word(i) >> 20 (12 bits).  Encoded block = W + 1 words:
  * no digit equals 0xFFF: words unchanged, extra word 0                                      (GF.md:81 "keep input data intact")
  * otherwise extra word 1 and the digit string becomes: for the j-th of the k digits equal to 0xFFF, in order, an entry
    `its position | (1 << 10 if another one follows)`; then the W - k other digits in order                 (GF.md:82-86)
  * the low 20 bits of every word stay in place; the extra bit travels as one more source word                (GF.md:101-103)
Every encoded word is <= 0xFFEFFFFF < P = 0xFFF00001."""
import numpy as np

P = 0xFFF00001


def bytes_to_gfp(words: np.ndarray) -> np.ndarray:
    """words: uint32 [n_blocks, W] -> uint32 [n_blocks, W + 1]"""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    n, W = words.shape
    assert W <= 1024
    out = np.zeros((n, W + 1), dtype=np.uint32)
    for b in range(n):
        w = words[b]
        d = (w >> 20).astype(np.int64)
        pos = np.flatnonzero(d == 0xFFF)
        if len(pos) == 0:
            out[b, :W] = w
            continue
        k = len(pos)
        entries = [int(p) | ((1 << 10) if j + 1 < k else 0) for j, p in enumerate(pos)]
        rest = [int(x) for x in d[d != 0xFFF]]
        nd = np.array(entries + rest, dtype=np.uint32)
        out[b, :W] = (nd << np.uint32(20)) | (w & np.uint32(0xFFFFF))
        out[b, W] = 1
    return out


def gfp_to_bytes(enc: np.ndarray) -> np.ndarray:
    """inverse: uint32 [n_blocks, W + 1] -> uint32 [n_blocks, W]"""
    enc = np.ascontiguousarray(enc, dtype=np.uint32)
    n, W1 = enc.shape
    W = W1 - 1
    out = np.zeros((n, W), dtype=np.uint32)
    for b in range(n):
        w = enc[b, :W]
        if enc[b, W] == 0:
            out[b] = w
            continue
        d = [int(x) for x in (w >> 20)]
        pos, j = [], 0
        while True:                                   # GF.md:83-85: index entries until the continuation flag is 0
            pos.append(d[j] & 0x3FF)
            more = d[j] & 0x400
            j += 1
            if not more:
                break
        rest = iter(d[j:])
        marks = set(pos)
        nd = np.array([0xFFF if i in marks else next(rest) for i in range(W)], dtype=np.uint32)
        out[b] = (nd << np.uint32(20)) | (w & np.uint32(0xFFFFF))
    return out
