// Warp-private tile schedule ("v9"): one warp owns a group of word columns of the tile for the whole tile.
//
// In ntt_tile.cuh's CTA-level schedule the exchange between the two register rounds of a transform goes through a
// block barrier, because the 32 rows a thread needs in round 1 were produced by threads of all eight warps.  Here
// a warp holds ALL R/32 row groups j of its columns (lane = jx * CQ + cq, CQ = 32 / (R/32) word pairs per warp), so
// the exchange is warp-private: __syncwarp instead of bar.sync, and the eight warps of a CTA drift apart freely
// (one is in its LDS phase while another multiplies).  The only CTA-wide events are the arrival of the tile
// (mbarrier, TMA) and "the last warp to finish writes the tile back" (a shared counter).
//
// Storage is the TMA tile itself, in its hardware-swizzled layout: natural row order, 64-byte rows with
// CU_TENSOR_MAP_SWIZZLE_64B (LR = 10), or column blocks of 128-byte rows with SWIZZLE_128B (LR <= 9).  A warp only
// ever touches the cells of its own columns, in three patterns, each with lanes <-> consecutive rows so that the
// swizzle spreads them over the banks (conflict-free for 128-byte rows, 2-way for 64-byte rows; tools/banksim.py):
//   natural   row = k*J + jx                       (first read: slot (j0<<5)|i holds input bitrev(slot), j0 = bitrev(jx);
//                                                   last write: output index jx + i*J)
//   exchange  row = phi(slot), written as slot (bitrev(jx)<<5)|i, read as slot jx | i<<(LR-5), with
//             phi(s) = s ^ (bitrev(s>>5) & 15): the low row bits are a bijection of jx for writers AND readers.
// Every address map above is GF(2)-linear in (jx, q2, i), so address = T(thread) ^ I(i): one thread constant per
// pattern (computed once per kernel) and one compile-time constant per access -- a single LOP3 in front of each LDS/STS.
#pragma once
#include "ntt_tile.cuh"

namespace fecc {

FECC_HD constexpr uint32_t bitrev_c(uint32_t x, uint32_t bits)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

struct WarpPos { uint32_t warp, jx, cq, q2; };
FECC_HD WarpPos warp_pos(uint32_t LR, uint32_t tid)
{
    const uint32_t cqbits = 10 - LR;                     // CQ = 32 / J, J = 2^(LR-5)
    WarpPos w;
    w.warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    w.cq = lane & ((1u << cqbits) - 1u);
    w.jx = lane >> cqbits;
    w.q2 = (w.warp << cqbits) | w.cq;                    // word-pair column inside the tile row
    return w;
}

// log2(bytes per row inside a column block): 64-byte rows for LR = 10, else blocks of 128-byte rows
FECC_HD constexpr uint32_t row_bytes_log(uint32_t LR) { return (16 - LR) < 7 ? (16 - LR) : 7; }

// uint2 index of cell (row, word pair q2) in the swizzled tile (the layout TMA writes and reads)
FECC_HD constexpr uint32_t cell8(uint32_t LR, uint32_t row, uint32_t q2)
{
    const uint32_t lrb = row_bytes_log(LR), lq = lrb - 3;
    const uint32_t blk = q2 >> lq, c2 = q2 & ((1u << lq) - 1u);
    const uint32_t chunk = (c2 >> 1) ^ (lrb == 6 ? ((row >> 1) & 3u) : (row & 7u));
    return (blk << (LR + lq)) | (row << lq) | (chunk << 1) | (c2 & 1u);
}

FECC_HD constexpr uint32_t xphi(uint32_t LR, uint32_t s)
{
    const uint32_t jbits = LR - 5, mbits = jbits < 4 ? jbits : 4;
    return s ^ (bitrev_c(s >> 5, jbits) & ((1u << mbits) - 1u));
}

// the three access patterns (uint2 index); all GF(2)-linear in (jx, q2, idx)
FECC_HD constexpr uint32_t addr_nat(uint32_t LR, uint32_t jx, uint32_t q2, uint32_t k) { return cell8(LR, (k << (LR - 5)) | jx, q2); }
FECC_HD constexpr uint32_t addr_xw (uint32_t LR, uint32_t jx, uint32_t q2, uint32_t i) { return cell8(LR, xphi(LR, (bitrev_c(jx, LR - 5) << 5) | i), q2); }
FECC_HD constexpr uint32_t addr_xr (uint32_t LR, uint32_t jx, uint32_t q2, uint32_t i) { return cell8(LR, xphi(LR, jx | (i << (LR - 5))), q2); }

struct WarpAddr { uint32_t nat, xw, xr; };              // thread parts T(jx, q2) of the three patterns
FECC_HD WarpAddr warp_addr(uint32_t LR, const WarpPos& w)
{
    WarpAddr a;
    a.nat = addr_nat(LR, w.jx, w.q2, 0);
    a.xw  = LR > 5 ? addr_xw(LR, w.jx, w.q2, 0) : 0;
    a.xr  = LR > 5 ? addr_xr(LR, w.jx, w.q2, 0) : 0;
    return a;
}

// first read of a tile: register i <- slot (j0<<5)|i = input element bitrev(slot) = natural row brev5(i)*J + jx
FECC_HD void warp_read_initial(uint32_t LR, const WarpAddr& a, const uint2* t2, RoundRegs& r)
{
#pragma unroll
    for (int i = 0; i < kRows; ++i) r.x[i] = t2[a.nat ^ addr_nat(LR, 0, 0, (uint32_t)brev5(i))];
}
// exchange: registers (slot (bitrev(jx)<<5)|i; after a fused step the slot i sits in register brev5(i)) -> cells
template <bool BREV>
FECC_HD void warp_xchg_write(uint32_t LR, const WarpAddr& a, uint2* t2, const RoundRegs& r)
{
#pragma unroll
    for (int i = 0; i < kRows; ++i) t2[a.xw ^ addr_xw(LR, 0, 0, (uint32_t)i)] = r.x[BREV ? brev5(i) : i];
}
// ... cells -> registers of the second round (slot jx | i<<(LR-5))
FECC_HD void warp_xchg_read(uint32_t LR, const WarpAddr& a, const uint2* t2, RoundRegs& r)
{
#pragma unroll
    for (int i = 0; i < kRows; ++i) r.x[i] = t2[a.xr ^ addr_xr(LR, 0, 0, (uint32_t)i)];
}
// last write of a tile: output element (natural row) i*J + jx, canonical residues, for the TMA store
template <bool BREV>
FECC_HD void warp_write_final(uint32_t LR, const WarpAddr& a, uint2* t2, const RoundRegs& r)
{
#pragma unroll
    for (int i = 0; i < kRows; ++i) t2[a.nat ^ addr_nat(LR, 0, 0, (uint32_t)i)] = canon2(r.x[BREV ? brev5(i) : i]);
}

FECC_HD bool warp_col_active(const PassParams& P, const WarpPos& w, uint32_t strip)
{
    return ((strip * (8192u >> P.log_r) + w.q2) >> 1) < P.s4;
}

// All the arithmetic of one tile between the first read and the last write (the caller provides the warp syncs
// through `sync`, a functor: __syncwarp on the device, nothing in the sequential CPU emulation where each phase is
// run for all threads before the next).  Split in phases so that both can drive it:
//   phase 0: read initial            phase 1: round 0 (+ prescale)       phase 2: exchange write
//   phase 3: exchange read           phase 4: round 1 [+ fused round 0 of the second transform]
//   phase 5: exchange write (fused)  phase 6: exchange read              phase 7: round 1 of the second transform
//   phase 8: final write
// Phases 2-7 exist only when LR > 5; 5-7 only when nxf == 2.
FECC_HD void warp_phase(const PassParams& P, int phase, uint32_t tid, uint32_t set, const WarpAddr& a, uint4* tile,
                        const uint4* tw0, const uint4* tw1, RoundRegs& r, uint32_t zero)
{
    const uint32_t LR = P.log_r;
    const WarpPos w = warp_pos(LR, tid);
    uint2* t2 = reinterpret_cast<uint2*>(tile);
    const bool two_rounds = LR > (uint32_t)kStages;
    const Xform xf0 = get_xf(P, 0);
    const bool plain0 = ((xf0.t0 + set * xf0.t1) & (gf::M - 1)) == 0;
    switch (phase) {
    case 0: warp_read_initial(LR, a, t2, r); break;
    case 1: {
        const RoundCtx c = make_round(LR, 0, bitrev(w.jx, LR - 5));
        if (plain0) round0_plain(r.x, c, tw0, P.prescale != 0, P.pw, P.pwhi, P.pwlo, zero);
        else { if (P.prescale) prescale_all(r.x, P.pw, P.pwhi, P.pwlo, zero); round_compute<false>(r.x, c, tw0, zero); }
        if (!two_rounds && P.nxf == 2) { const RoundCtx c1 = make_round(LR, 0, 0); round_compute<true>(r.x, c1, tw1, zero); }
        break; }
    case 2: warp_xchg_write<false>(LR, a, t2, r); break;
    case 3: warp_xchg_read(LR, a, t2, r); break;
    case 4: {
        const RoundCtx c = make_round(LR, 1, w.jx);
        round_compute<false>(r.x, c, tw0, zero);
        if (P.nxf == 2) { const RoundCtx c1 = make_round(LR, 0, 0); round_compute<true>(r.x, c1, tw1, zero); }
        break; }
    case 5: warp_xchg_write<true>(LR, a, t2, r); break;
    case 6: warp_xchg_read(LR, a, t2, r); break;
    case 7: { const RoundCtx c = make_round(LR, 1, w.jx); round_compute<false>(r.x, c, tw1, zero); break; }
    case 8:
        if (!two_rounds && P.nxf == 2) warp_write_final<true>(LR, a, t2, r);     // single fused round: slot i in register brev5(i)
        else                           warp_write_final<false>(LR, a, t2, r);
        break;
    }
}

} // namespace fecc
