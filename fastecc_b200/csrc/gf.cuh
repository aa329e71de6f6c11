// GF(0xFFF00001) arithmetic for the sm_100a NTT kernels (and a bit-identical host restatement used by the
// CPU emulation of the kernels and by the twiddle-table generator).
//
// Replaces, on the device, GF_Add / GF_Sub / GF_Mul of the reference (GF(p).cpp:37-48, 110-127).  The reference's
// 32-bit Barrett (GF_Mul32) needs a 64-bit conditional subtract per product; here every twiddle w is a constant
// known on the host, so we use an exact Barrett/Shoup product with a precomputed 64-bit quotient
//      W = floor(w * 2^64 / P) = (Whi:Wlo)
//      q = floor(b * W / 2^64) = hi32( b*Whi + hi32(b*Wlo) )            (2 x IMAD.HI)
//      v = lo32(b*w) - lo32(q*P) = b*w + q*(2^32-P)   (mod 2^32)         (2 x IMAD)
// q equals floor(b*w/P) exactly unless b*w is a multiple of P, in which case it may be one less; so
//      v == b*w mod P   with v in [0, P]          for ANY 32-bit b  (b need not be reduced).
// No conditional subtract.  Values between butterflies are kept "lazy" in [0, 2^32) (congruent mod P):
//      addl(a, v):  a in [0,2^32), v in [0,P]  ->  a+v mod P  in [0,2^32)     (add, then +(2^32-P) on carry)
//      subl(a, v):  a in [0,2^32), v in [0,P]  ->  a-v mod P  in [0,2^32)     (sub, then -(2^32-P) on borrow)
// which are closed (proof in DESIGN.md section 4); canon() maps a lazy value to the canonical residue [0,P).
// The carry idioms below compile to IADD3 Rd,Pc + @Pc VIADD (2 SASS instructions); both were validated on a B200
// against 64-bit arithmetic (profiles/r01_ubench_int_pipes.log, "selftest" lines).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GF_HD __host__ __device__ __forceinline__
#else
#define GF_HD inline
#endif

namespace gf {

constexpr uint32_t P  = 0xFFF00001u;
constexpr uint32_t C  = 0x000FFFFFu;          // 2^32 - P
constexpr uint32_t GEN = 19;                  // generator used by GF_Root (GF(p).cpp:272)
constexpr uint32_t LOG_M = 20;                // largest power-of-two order: P-1 = 2^20 * 4095
constexpr uint32_t M = 1u << LOG_M;

struct Tw { uint32_t w, whi, wlo, wm; };      // one twiddle: w, floor(w*2^64/P), and w*2^32 mod P (Montgomery form); 16 bytes

// `zero` must be a register holding 0 that the compiler cannot constant-fold (see opaque_zero()): it becomes the
// high half of the 64-bit addend of the second IMAD.HI, which saves ptxas from re-materialising a zero register
// next to `t` for every product (one extra FMA-pipe instruction per butterfly otherwise).
GF_HD uint32_t mul(uint32_t b, uint32_t w, uint32_t whi, uint32_t wlo, uint32_t zero = 0)
{
#if defined(__CUDA_ARCH__)
    uint32_t t = __umulhi(b, wlo);
    uint64_t addend = ((uint64_t)zero << 32) | t;
    uint32_t q = (uint32_t)(((uint64_t)b * whi + addend) >> 32);
#else
    (void)zero;
    uint32_t t = (uint32_t)(((uint64_t)b * wlo) >> 32);
    uint32_t q = (uint32_t)(((uint64_t)b * whi + t) >> 32);
#endif
    return q * C + b * w;
}

// The same product through a Montgomery reduction (R = 2^32): wm = w*2^32 mod P.  P^-1 mod 2^32 = 1 + 2^20, so the
// reduction factor m = lo*(1 + 2^20) is one LEA on the ALU pipe; T - m*P is divisible by 2^32 and
// (T - m*P)/2^32 = hi(T) - hi(m*P) lies in (-P, P).  IMAD.WIDE + IMAD.HI + 3 ALU instructions: fewer cycles on the
// integer-multiply pipe than mul() (2 x IMAD.HI + 2 x IMAD), more on the ALU pipe -- the kernels use mul() for one
// word of a pair and mul_mont() for the other to load both pipes evenly.  Result in [0, P).
GF_HD uint32_t mul_mont(uint32_t b, uint32_t wm)
{
    const uint64_t T = (uint64_t)b * wm;
    const uint32_t lo = (uint32_t)T, hi = (uint32_t)(T >> 32);
    const uint32_t m = lo + (lo << 20);
#if defined(__CUDA_ARCH__)
    const uint32_t h2 = __umulhi(m, P);
    uint32_t r;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tsub.cc.u32 %0, %1, %2;\n\taddc.u32 c, 0, 0;\n\tsetp.eq.u32 q, c, 0;\n\t@q add.u32 %0, %0, 0xFFF00001;\n\t}"
        : "=r"(r) : "r"(hi), "r"(h2));
    return r;
#else
    const uint32_t h2 = (uint32_t)(((uint64_t)m * P) >> 32);
    const uint32_t r = hi - h2;
    return hi < h2 ? r + P : r;
#endif
}

GF_HD uint32_t addl(uint32_t a, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    uint32_t s;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tadd.cc.u32 %0, %1, %2;\n\taddc.u32 c, 0, 0;\n\tsetp.ne.u32 q, c, 0;\n\t@q add.u32 %0, %0, 0xFFFFF;\n\t}"
        : "=r"(s) : "r"(a), "r"(v));
    return s;
#else
    uint32_t s = a + v; return s < a ? s + C : s;
#endif
}

GF_HD uint32_t subl(uint32_t a, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    // sub.cc leaves CF = NOT borrow for a following addc on this toolchain/hardware (checked on B200: the
    // "subfix1" line of the self-test); so "c == 0" means a borrow happened.
    uint32_t s;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tsub.cc.u32 %0, %1, %2;\n\taddc.u32 c, 0, 0;\n\tsetp.eq.u32 q, c, 0;\n\t@q sub.u32 %0, %0, 0xFFFFF;\n\t}"
        : "=r"(s) : "r"(a), "r"(v));
    return s;
#else
    uint32_t s = a - v; return a < v ? s - C : s;
#endif
}

GF_HD uint32_t opaque_zero()
{
#if defined(__CUDA_ARCH__)
    uint32_t z; asm volatile("mov.u32 %0, 0;" : "=r"(z)); return z;
#else
    return 0;
#endif
}

GF_HD uint32_t canon(uint32_t x) { return x >= P ? x - P : x; }

// ---- host-only helpers (table generation, plan constants) --------------------------------------------------------
inline uint32_t mulmod(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
inline uint32_t powmod(uint32_t x, uint64_t n) { uint32_t r = 1; for (; n; n >>= 1) { if (n & 1) r = mulmod(r, x); x = mulmod(x, x); } return r; }
inline uint32_t root(uint32_t n) { return powmod(GEN, (P - 1) / n); }        // GF_Root, GF(p).cpp:267-276
inline uint32_t inv(uint32_t x)  { return powmod(x, P - 2); }                // GF_Inv,  GF(p).cpp:293-297
inline Tw make_tw(uint32_t w)            // {w, floor(w*2^64/P)} by two 64/32 long-division steps
{
    const uint64_t n1 = (uint64_t)w << 32;
    const uint64_t whi = n1 / P, rem = n1 % P;
    const uint64_t wlo = (rem << 32) / P;
    Tw t; t.w = w; t.whi = (uint32_t)whi; t.wlo = (uint32_t)wlo; t.wm = (uint32_t)(n1 % P); return t;
}

} // namespace gf
