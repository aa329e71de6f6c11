"""Bank-conflict count of the CTA-level tile schedule (ntt_tile.cuh): for every shared-memory access instruction of a
tile (round_read / round_write_tile, brev or identity placement) the number of extra wavefronts a warp needs.
LDS.64 / STS.64 are served per half-warp: 16 lanes x 8 B = 128 B = all 32 banks when conflict-free.
Optional layout: the popcount-parity row swap (row ^= parity(row >> 1)) used by the cp.async version of the kernel."""
import sys

def brev(x, bits):
    r = 0
    for i in range(bits): r |= ((x >> i) & 1) << (bits - 1 - i)
    return r

def parity(x): return bin(x).count("1") & 1

def addresses(LR, k, use_brev, tid):
    """uint2 indices of the 32 accesses of thread tid in round k (FECC_SLOT_LOOP)"""
    q2log = 13 - LR
    q2, j = tid & ((1 << q2log) - 1), tid >> q2log
    lb = 0 if k == 0 else LR - 5
    jlow = j & ((1 << lb) - 1)
    jbase = ((j >> lb) << (lb + 5)) | jlow
    if use_brev:
        a = (brev(jbase, LR) << q2log) | q2
        step = 1 << (LR - 5 - lb + q2log)
    else:
        a = (jbase << q2log) | q2
        step = 1 << (lb + q2log)
    return [a + i * step for i in range(32)]

def extra_wavefronts(LR, k, use_brev, swap):
    q2log = 13 - LR
    extra = total = 0
    for warp in range(8):
        per_thread = [addresses(LR, k, use_brev, warp * 32 + l) for l in range(32)]
        for i in range(32):
            for half in (range(0, 16), range(16, 32)):
                banks = {}
                for l in half:
                    a = per_thread[l][i]
                    row, q2 = a >> q2log, a & ((1 << q2log) - 1)
                    if swap: row ^= parity(row >> 1)
                    byte = (row << (q2log + 3)) + q2 * 8
                    banks.setdefault((byte >> 3) & 15, set()).add(byte)
                extra += max(len(v) for v in banks.values()) - 1
                total += 1
    return extra, total

if __name__ == "__main__":
    for LR in (10, 9, 8):
        for swap in (False, True):
            parts = []
            for name, k, b in (("first transform round 0", 0, True), ("first transform round 1", 1, True), ("second transform round 0", 0, False), ("second transform round 1", 1, False)):
                e, t = extra_wavefronts(LR, k, b, swap)
                parts.append("%s: %d/%d" % (name, e, t))
            print("LR=%d %s  extra wavefronts / half-warp accesses:  %s" % (LR, "parity-swapped rows" if swap else "natural rows (TMA)  ", "; ".join(parts)))
