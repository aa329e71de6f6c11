"""Bank-conflict simulator for the warp-private exchange design (v9): for every access pattern of a tile, the max number
of distinct 8-byte cells that fall into the same bank pair within a half-warp (LDS.64/STS.64 are served per half-warp)."""
import itertools

def brev(x, bits):
    r = 0
    for i in range(bits): r |= ((x >> i) & 1) << (bits - 1 - i)
    return r

def cell_byte(LR, row, q2):
    """byte offset of (row, word pair q2) in the TMA tile: column blocks of <=128 B rows, TMA swizzle (SW64 for 64 B rows, SW128 for 128 B)."""
    wt_bytes = (16384 >> LR) * 4
    rb = min(wt_bytes, 128)                       # bytes per row inside a block
    per_row_q2 = rb // 8
    blk, c2 = divmod(q2, per_row_q2)
    chunk, half = c2 >> 1, c2 & 1
    R = 1 << LR
    if rb == 64:  chunk ^= (row >> 1) & 3
    else:         chunk ^= row & 7
    return blk * R * rb + row * rb + chunk * 16 + half * 8

def conflicts(addrs):
    """addrs: 32 byte addresses (lane order). returns max multiplicity of bank-pairs over the two half-warps (distinct addresses only)."""
    worst = 0
    for h in (addrs[:16], addrs[16:]):
        cnt = {}
        for a in set(h): cnt.setdefault((a >> 3) & 15, set()).add(a)
        worst = max(worst, max(len(v) for v in cnt.values()))
    return worst

def phi(LR, s):
    J = 1 << (LR - 5)
    mbits = min(4, LR - 5)
    m = (1 << mbits) - 1
    return s ^ (brev(s >> 5, LR - 5) & m) if LR > 5 else s

for LR in (10, 9, 8, 7, 6):
    J = 1 << (LR - 5); CQ = 32 // J
    res = {}
    for w in (0, 3):
        lanes = [(l // CQ, w * CQ + l % CQ) for l in range(32)]           # (jx, q2)
        # initial read / final write: rows k*J + jx
        res['natural'] = max(res.get('natural', 0), max(conflicts([cell_byte(LR, k * J + jx, q2) for jx, q2 in lanes]) for k in range(32)))
        # exchange writer: thread jw = brev(jx), slots (jw<<5)|i
        res['xw'] = max(res.get('xw', 0), max(conflicts([cell_byte(LR, phi(LR, (brev(jx, LR - 5) << 5) | i), q2) for jx, q2 in lanes]) for i in range(32)))
        # exchange reader: thread jr = jx, slots jr | i << (LR-5)
        res['xr'] = max(res.get('xr', 0), max(conflicts([cell_byte(LR, phi(LR, jx | (i << (LR - 5))), q2) for jx, q2 in lanes]) for i in range(32)))
    # bijection check of phi
    assert sorted(phi(LR, s) for s in range(1 << LR)) == list(range(1 << LR))
    print("LR=%d J=%d CQ=%d  worst conflict degree: natural %d, exchange write %d, exchange read %d" % (LR, J, CQ, res['natural'], res['xw'], res['xr']))
