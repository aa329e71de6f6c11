// Tile-level decimation-in-time NTT over GF(0xFFF00001): the code every pass kernel runs per 64 KiB tile.
//
// This header is the B200 counterpart of the reference's inner loops
//     IterativeNTT_Steps  ntt.cpp:251-284   (radix-2 butterflies across blocks, inner loop over the SIZE words)
//     revbin_permute      ntt.cpp:292-309   (here: bit-reversed *addressing* when a tile is loaded, no data pass)
//     MFA twiddle loop    ntt.cpp:421-431   (here: folded into the butterfly twiddles of the next pass, see below)
//     scaling loop        RS.cpp:51-59      (here: folded into butterfly twiddles + one uniform pre-scale)
// It is written as plain `__host__ __device__` code over an abstract "thread id" so that exactly the same
// index/twiddle logic can be executed thread-by-thread on the CPU (tests/emulate_tile.cpp) before it ever
// touches a GPU.
//
// Geometry.  A tile is R = 2^LR rows (blocks) x WT words, R*WT = 16384 words = 4096 16-byte chunks (64 KiB of
// shared memory).  256 threads; thread (j, q) owns chunk column q (4 consecutive words of every row: the
// word dimension is a pure batch dimension, SURVEY 7 "hard part 2") and, in every round, 16 rows.  With
// Q = WT/4 chunks per row:  q = tid % Q,  j = tid / Q  in [0, R/16).
//
// Rounds.  A size-R DIT transform runs ceil(LR/4) rounds; round k handles index bits [lb, lb+4), lb = min(4k, LR-4),
// executing the radix-2 stages for bits >= 4k only.  In a round a thread holds the 16 slots
//     r_i = ((j >> lb) << (lb+4)) | (i << lb) | (j & (2^lb - 1)),   i = 0..15
// in registers (16 x uint4), performs up to 4 stages x 8 butterflies x 4 words, and writes the slots back in
// place; one block barrier separates rounds.  Slot r of a DIT transform initially holds input element
// bitrev_LR(r) and finally holds output element r.
//
// Twiddles.  All twiddles are powers of g = GF_Root(2^20) (GF(p).cpp:267-276).  A transform is described by two
// exponents (mod 2^20):  zeta = g^z is the primitive R-th root used by this tile, theta = g^t an *input twist*:
// the transform computes  X[k] = sum_n x[n] * theta^n * zeta^(n k).  A twisted DIT costs exactly as much as a
// plain one: the stage that pairs slots differing in bit b uses, for the pair whose low b bits are n,
//     T[2^b + n] = g^( 2^(LR-1-b) * (z*n + t) )                                   (heap-ordered table, R-1 entries)
// (derivation in DESIGN.md section 3).  Four-step twiddles (ntt.cpp:421-431) and the RS scaling root_2N^i
// (RS.cpp:54) are input twists of the following pass, so they cost no multiplications at all.
#pragma once
#include "gf.cuh"

#if defined(__CUDACC__)
#define FECC_HD __host__ __device__ __forceinline__
#else
#define FECC_HD inline
#endif

namespace fecc {

constexpr int kThreads        = 256;
constexpr int kTileChunks     = 4096;            // 16-byte chunks per tile
constexpr int kTileBytes      = kTileChunks * 16;
constexpr int kMaxLogR        = 10;
constexpr int kMinLogR        = 4;

struct Xform { uint32_t z, t0, t1; };            // tile root g^z; input twist g^(t0 + set*t1)   (exponents mod 2^20)

struct PassParams {
    const uint32_t* src;
    uint32_t*       dst;
    const uint4*    tw;                          // g^e for e in [0, 2^20): {w, Whi, Wlo, 0}
    uint32_t pitch4;                             // row pitch in 16-byte chunks
    uint32_t s4;                                 // valid 16-byte chunks per row (ceil(SIZE/4))
    uint32_t log_r;                              // log2(rows per tile)
    uint32_t nsets;                              // number of independent row sets
    uint32_t nstrips;                            // ceil(s4 / Q)
    uint32_t strips_per_item;                    // consecutive strips of one set handled by one CTA visit
    uint32_t src_set_stride, src_row_stride, dst_set_stride, dst_row_stride;             // in rows
    uint32_t nxf;                                // 1, or 2 = two transforms back to back on the same tile
    Xform    xf[2];
    uint32_t prescale;                           // multiply every input word by the constant below (1/N)
    uint32_t pw, pwhi, pwlo;
    uint32_t canonical_out;                      // reduce stored words to [0,P)
    uint32_t parity;                             // Q==4 layout: swap the two rows of a pair when popcount(row>>1) is odd
};

FECC_HD uint32_t bitrev(uint32_t x, uint32_t bits)
{
#if defined(__CUDA_ARCH__)
    return __brev(x) >> (32 - bits);
#else
    uint32_t r = 0; for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i); return r;
#endif
}
FECC_HD uint32_t popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

// chunk index (in uint4 units) of (physical row p, chunk q) inside the tile
FECC_HD uint32_t tile_chunk(uint32_t p, uint32_t q, uint32_t qlog, uint32_t parity)
{
    if (parity) p ^= popc32(p >> 1) & 1u;
    return (p << qlog) | q;
}

// Rounds of a size-2^LR transform: nr = ceil(LR/4).  Round 0 executes the first rem = LR - 4(nr-1) stages (bits
// [0, rem)) with bits 0..3 in-thread; round k >= 1 executes the four stages of bits [lb, lb+4), lb = rem + 4(k-1).
// Putting the short round first leaves a full four-stage round at the end of the tile, which is the compute that
// hides the next tile's cp.async loads (ntt_pass.cu).
FECC_HD uint32_t num_rounds(uint32_t LR) { return (LR + 3) >> 2; }
FECC_HD void round_bits(uint32_t LR, uint32_t k, uint32_t& lb, uint32_t& blo, uint32_t& bhi)
{
    const uint32_t rem = LR - 4 * (num_rounds(LR) - 1);
    if (k == 0) { lb = 0; blo = 0; bhi = rem; }
    else        { lb = rem + 4 * (k - 1); blo = lb; bhi = lb + 4; }
}

// exponent of heap entry idx (1 <= idx < R) of the stage table
FECC_HD uint32_t table_exponent(uint32_t idx, uint32_t LR, uint32_t z, uint32_t t)
{
#if defined(__CUDA_ARCH__)
    uint32_t b = 31 - __clz(idx);
#else
    uint32_t b = 31 - (uint32_t)__builtin_clz(idx);
#endif
    uint32_t n = idx - (1u << b);
    return ((z * n + t) << (LR - 1 - b)) & (gf::M - 1);
}

// One butterfly on 4 words:  (a, b) <- (a + w*b, a - w*b)
FECC_HD void bfly4(uint4& a, uint4& b, const uint4& w, uint32_t zero)
{
    uint32_t v;
    v = gf::mul(b.x, w.x, w.y, w.z, zero); b.x = gf::subl(a.x, v); a.x = gf::addl(a.x, v);
    v = gf::mul(b.y, w.x, w.y, w.z, zero); b.y = gf::subl(a.y, v); a.y = gf::addl(a.y, v);
    v = gf::mul(b.z, w.x, w.y, w.z, zero); b.z = gf::subl(a.z, v); a.z = gf::addl(a.z, v);
    v = gf::mul(b.w, w.x, w.y, w.z, zero); b.w = gf::subl(a.w, v); a.w = gf::addl(a.w, v);
}

FECC_HD uint4 canon4(uint4 v) { v.x = gf::canon(v.x); v.y = gf::canon(v.y); v.z = gf::canon(v.z); v.w = gf::canon(v.w); return v; }

// Slot bookkeeping of one thread in one round.
struct RoundCtx {
    uint32_t lb, blo, bhi;    // bits [lb, lb+4) are in-thread; the stages of bits [blo, bhi) are executed
    uint32_t jbase;           // slot index with the in-thread bits zero
    uint32_t jlow;            // low lb bits of j (= low lb bits of every slot of this thread)
};

FECC_HD RoundCtx make_round(uint32_t LR, uint32_t k, uint32_t j)
{
    RoundCtx c;
    round_bits(LR, k, c.lb, c.blo, c.bhi);
    c.jlow  = j & ((1u << c.lb) - 1u);
    c.jbase = ((j >> c.lb) << (c.lb + 4)) | c.jlow;
    return c;
}

// The register-resident part of a round: up to four radix-2 DIT stages on the thread's 16 slots x 4 words.
// tw points at the heap-ordered stage table of the current transform (shared memory on the device).
FECC_HD void round_compute(uint4 (&x)[16], const RoundCtx& c, const uint4* tw, uint32_t zero)
{
    const uint32_t sb = 1u << c.lb;                     // table stride between consecutive in-thread twiddles
    const uint4* tb = tw + c.jlow;
#pragma unroll
    for (int beta = 0; beta < 4; ++beta) {
        const uint32_t b = c.lb + beta;
        if (b >= c.blo && b < c.bhi) {
            const uint4* twp = tb + (sb << beta);       // heap entry 2^b + jlow (+ m * 2^lb)
#pragma unroll
            for (int m = 0; m < (1 << beta); ++m) {
                const uint4 w = *twp; twp += sb;
#pragma unroll
                for (int hi = 0; hi < (8 >> beta); ++hi) {
                    const int i0 = (hi << (beta + 1)) | m;
                    const int i1 = i0 | (1 << beta);
                    bfly4(x[i0], x[i1], w, zero);
                }
            }
        }
    }
}

// Round 0 of a PLAIN transform (no input twist): the twiddle of every pair whose low bits are zero is 1, which is
// known at compile time because lb = 0 makes the table index depend on the in-thread index only.  Those
// butterflies need no product: the b operand is merely brought to [0,P) (two ALU instructions) so that the lazy
// add/sub stay closed.  The uniform pre-scale by 1/N (RS.cpp:51,54) is folded into stage 0, whose 8 butterflies
// then cost two products each instead of one product plus two pre-scale products.
FECC_HD void round0_plain(uint4 (&x)[16], const RoundCtx& c, const uint4* tw, bool PRESCALE, uint32_t pw, uint32_t pwhi, uint32_t pwlo, uint32_t zero)
{
    const uint4 cw = {pw, pwhi, pwlo, 0};
#pragma unroll
    for (int hi = 0; hi < 8; ++hi) {                       // stage 0: all twiddles are 1
        uint4& a = x[2 * hi]; uint4& b = x[2 * hi + 1];
        if (PRESCALE) {
            a.x = gf::mul(a.x, cw.x, cw.y, cw.z, zero); a.y = gf::mul(a.y, cw.x, cw.y, cw.z, zero);
            a.z = gf::mul(a.z, cw.x, cw.y, cw.z, zero); a.w = gf::mul(a.w, cw.x, cw.y, cw.z, zero);
            b.x = gf::mul(b.x, cw.x, cw.y, cw.z, zero); b.y = gf::mul(b.y, cw.x, cw.y, cw.z, zero);
            b.z = gf::mul(b.z, cw.x, cw.y, cw.z, zero); b.w = gf::mul(b.w, cw.x, cw.y, cw.z, zero);
        } else {
            b = canon4(b);
        }
        uint32_t t;
        t = b.x; b.x = gf::subl(a.x, t); a.x = gf::addl(a.x, t);
        t = b.y; b.y = gf::subl(a.y, t); a.y = gf::addl(a.y, t);
        t = b.z; b.z = gf::subl(a.z, t); a.z = gf::addl(a.z, t);
        t = b.w; b.w = gf::subl(a.w, t); a.w = gf::addl(a.w, t);
    }
#pragma unroll
    for (int beta = 1; beta < 4; ++beta) {
        if ((uint32_t)beta < c.bhi) {
            const uint4* twb = tw + (1u << beta);
#pragma unroll
            for (int m = 0; m < (1 << beta); ++m) {
                uint4 w = {0, 0, 0, 0};
                if (m) w = twb[m];
#pragma unroll
                for (int hi = 0; hi < (8 >> beta); ++hi) {
                    const int i0 = (hi << (beta + 1)) | m;
                    const int i1 = i0 | (1 << beta);
                    if (m) {
                        bfly4(x[i0], x[i1], w, zero);
                    } else {
                        uint4& a = x[i0]; uint4 b = canon4(x[i1]);
                        x[i1].x = gf::subl(a.x, b.x); a.x = gf::addl(a.x, b.x);
                        x[i1].y = gf::subl(a.y, b.y); a.y = gf::addl(a.y, b.y);
                        x[i1].z = gf::subl(a.z, b.z); a.z = gf::addl(a.z, b.z);
                        x[i1].w = gf::subl(a.w, b.w); a.w = gf::addl(a.w, b.w);
                    }
                }
            }
        }
    }
}

// Compile-time helpers for the unrolled slot loops
FECC_HD constexpr uint32_t par4(int i)  { return (0x6996u >> i) & 1u; }                                   // popcount(i) & 1
FECC_HD constexpr int      brev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

FECC_HD void prescale16(uint4 (&x)[16], uint32_t w, uint32_t whi, uint32_t wlo, uint32_t zero)
{
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i].x = gf::mul(x[i].x, w, whi, wlo, zero);
        x[i].y = gf::mul(x[i].y, w, whi, wlo, zero);
        x[i].z = gf::mul(x[i].z, w, whi, wlo, zero);
        x[i].w = gf::mul(x[i].w, w, whi, wlo, zero);
    }
}

} // namespace fecc

// ---------------------------------------------------------------------------------------------------------------
// Per-thread steps of a tile.  The kernel (ntt_pass.cu) runs them with block barriers in between; the CPU
// emulation (tests/emulate_tile.cu) runs each step for tid = 0..255 in turn, which is equivalent because within a
// step a thread only touches its own slots.
// ---------------------------------------------------------------------------------------------------------------
namespace fecc {

// P.xf[xfi] with a runtime xfi would make the compiler copy the kernel parameters to local memory
FECC_HD Xform get_xf(const PassParams& P, uint32_t xfi)
{
    Xform x;
    x.z  = xfi ? P.xf[1].z  : P.xf[0].z;
    x.t0 = xfi ? P.xf[1].t0 : P.xf[0].t0;
    x.t1 = xfi ? P.xf[1].t1 : P.xf[0].t1;
    return x;
}

struct ThreadPos { uint32_t q, j, qlog, Q; };
FECC_HD ThreadPos thread_pos(const PassParams& P, uint32_t tid)
{
    ThreadPos t; t.qlog = 12 - P.log_r; t.Q = 1u << t.qlog; t.q = tid & (t.Q - 1); t.j = tid >> t.qlog; return t;
}

FECC_HD void copy16(uint4* dst_smem, const uint4* src_gmem)
{
#if defined(__CUDA_ARCH__)
    uint32_t s = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(src_gmem));
#else
    *dst_smem = *src_gmem;
#endif
}

// Gather the heap-ordered stage table of transform xfi for row set `set` from the global power table (cp.async).
FECC_HD void build_table(const PassParams& P, uint32_t xfi, uint32_t set, uint32_t tid, uint4* tw_s)
{
    const uint32_t R = 1u << P.log_r;
    const Xform xf = get_xf(P, xfi);
    const uint32_t z = xf.z;
    const uint32_t t = (xf.t0 + set * xf.t1) & (gf::M - 1);
    for (uint32_t idx = tid; idx < R; idx += kThreads)
        if (idx) copy16(tw_s + idx, P.tw + table_exponent(idx, P.log_r, z, t));
}

// Issue the (asynchronous) loads of tile (set, strip).  Thread tid moves the 16 chunks c = tid + 256*m: tile row
// p = p0 + m*2^(LR-4), whose source row is bitrev(p) = bitrev(p0) + brev4(m) -- sixteen consecutive source rows,
// walked with one pointer increment each.
FECC_HD void load_tile(const PassParams& P, uint32_t set, uint32_t strip, uint32_t tid, uint4* tile)
{
    const uint32_t LR = P.log_r, qlog = 12 - LR, Q = 1u << qlog;
    const uint32_t p0 = tid >> qlog, qq = tid & (Q - 1);
    const uint32_t gcol = strip * Q + qq;
    if (gcol >= P.s4) return;
    const uint32_t row0 = set * P.src_set_stride + bitrev(p0, LR) * P.src_row_stride;
    const uint4* g = reinterpret_cast<const uint4*>(P.src) + ((size_t)row0 * P.pitch4 + gcol);
    const size_t gstep = (size_t)P.src_row_stride * P.pitch4;
    uint32_t cbase = (p0 << qlog) | qq;
    if (P.parity) cbase ^= (popc32(p0 >> 1) & 1u) << qlog;
    const uint32_t pbit = P.parity ? (1u << qlog) : 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {                          // k = brev4(m): k-th consecutive source row
        copy16(tile + ((cbase + 256u * (uint32_t)brev4(k)) ^ (par4(k) ? pbit : 0u)), g);
        g += gstep;
    }
}

struct RoundRegs { uint4 x[16]; };

FECC_HD bool thread_active(const PassParams& P, uint32_t tid, uint32_t strip)
{
    const ThreadPos tp = thread_pos(P, tid);
    return strip * tp.Q + tp.q < P.s4;                          // partial last strip: column chunk beyond the row
}

// Tile chunk of the thread's slot i in round c:  physical row = slot (first transform) or bitrev(slot) (second
// transform of a fused tile), rows of a pair swapped by popcount parity when P.parity.  All variants reduce to
// "base + i*step" (or base ^ const) so that a round spends one ALU instruction per 16-byte access.
#define FECC_SLOT_LOOP(ACCESS)                                                                                     \
    if (P.parity) {                    /* identity placement only: plan.h never combines parity with a fused tile */ \
        if (c.lb == 0) {                                                                                           \
            const uint32_t B = ((tp.j << 6) | tp.q) ^ ((popc32(tp.j) & 1u) << 2);                                  \
            _Pragma("unroll") for (int i = 0; i < 16; ++i) { const uint32_t a = B ^ (uint32_t)((i ^ (int)par4(i >> 1)) << 2); ACCESS(i, a); } \
        } else {                                                                                                   \
            uint32_t a0 = ((c.jbase ^ (popc32(c.jbase >> 1) & 1u)) << 2) | tp.q;                                   \
            const uint32_t step = 4u << c.lb;                                                                      \
            _Pragma("unroll") for (int i = 0; i < 16; ++i) { const uint32_t a = a0 ^ (par4(i) << 2); ACCESS(i, a); a0 += step; } \
        }                                                                                                          \
    } else if (xfi) {                                                                                              \
        uint32_t a = (bitrev(c.jbase, LR) << tp.qlog) | tp.q;                                                      \
        const uint32_t step = 1u << (LR - 4 - c.lb + tp.qlog);                                                     \
        _Pragma("unroll") for (int k = 0; k < 16; ++k) { ACCESS(brev4(k), a); a += step; }                         \
    } else {                                                                                                       \
        uint32_t a = (c.jbase << tp.qlog) | tp.q;                                                                  \
        const uint32_t step = 1u << (c.lb + tp.qlog);                                                              \
        _Pragma("unroll") for (int i = 0; i < 16; ++i) { ACCESS(i, a); a += step; }                                \
    }

// (a) read the thread's 16 slots of round k of transform xfi
FECC_HD void round_read(const PassParams& P, uint32_t xfi, uint32_t k, uint32_t tid, const uint4* tile, RoundRegs& r)
{
    const uint32_t LR = P.log_r;
    const ThreadPos tp = thread_pos(P, tid);
    const RoundCtx c = make_round(LR, k, tp.j);
#define FECC_RD(I, A) r.x[I] = tile[A]
    FECC_SLOT_LOOP(FECC_RD)
#undef FECC_RD
}

// (b) the butterflies
FECC_HD void round_math(const PassParams& P, uint32_t xfi, uint32_t k, uint32_t tid, uint32_t set, const uint4* tw_s, RoundRegs& r, uint32_t zero)
{
    const ThreadPos tp = thread_pos(P, tid);
    const RoundCtx c = make_round(P.log_r, k, tp.j);
    const bool first = (xfi == 0 && k == 0);
    const Xform xf = get_xf(P, xfi);
    const bool plain = ((xf.t0 + set * xf.t1) & (gf::M - 1)) == 0;
    if (k == 0 && plain) {
        round0_plain(r.x, c, tw_s, first && P.prescale, P.pw, P.pwhi, P.pwlo, zero);
    } else {
        if (first && P.prescale) prescale16(r.x, P.pw, P.pwhi, P.pwlo, zero);
        round_compute(r.x, c, tw_s, zero);
    }
}

// (c) write back in place, or (last round of the last transform) store output element r to its global row
FECC_HD void round_write_tile(const PassParams& P, uint32_t xfi, uint32_t k, uint32_t tid, uint4* tile, const RoundRegs& r)
{
    const uint32_t LR = P.log_r;
    const ThreadPos tp = thread_pos(P, tid);
    const RoundCtx c = make_round(LR, k, tp.j);
#define FECC_WR(I, A) tile[A] = r.x[I]
    FECC_SLOT_LOOP(FECC_WR)
#undef FECC_WR
}
FECC_HD void round_write_global(const PassParams& P, uint32_t k, uint32_t tid, uint32_t set, uint32_t strip, const RoundRegs& r)
{
    const ThreadPos tp = thread_pos(P, tid);
    const RoundCtx c = make_round(P.log_r, k, tp.j);
    const uint32_t gcol = strip * tp.Q + tp.q;
    const uint32_t row0 = set * P.dst_set_stride + c.jbase * P.dst_row_stride;
    uint4* g = reinterpret_cast<uint4*>(P.dst) + ((size_t)row0 * P.pitch4 + gcol);
    const size_t gstep = ((size_t)P.dst_row_stride * P.pitch4) << c.lb;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        *g = P.canonical_out ? canon4(r.x[i]) : r.x[i];
        g += gstep;
    }
}

} // namespace fecc
