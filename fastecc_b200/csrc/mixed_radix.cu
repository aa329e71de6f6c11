// Orders N = 3 * 2^k and 9 * 2^k (P - 1 = 2^20 * 3^2 * 5 * 7 * 13), so that block counts need not be rounded up to a power of
// two (SURVEY 8f rank 4 tail; the reference has the order-3 / order-9 codelets NTT3, NTT9 -- ntt.cpp:26-46, 114-146 -- and their
// composition with the power-of-two transforms on its roadmap, README.md:176).  Cooley-Tukey with N = r * M, r in {3, 9}, M = 2^k,
// input index n = n1 + r*n2, output index k = k2 + M*k1:
//     X[k2 + M*k1] = sum_{n1 < r}  w_r^(n1*k1)  *  [ w_N^(n1*k2) * Y_n1[k2] ],      Y_n1 = M-point NTT of x[n1 + r*n2] over n2
// with w_N = GF_Root(N), w_N^r = GF_Root(M) (the root of the power-of-two kernels) and w_N^M = GF_Root(r).  Step 1 -- the r
// transforms Y_n1 -- is the hot-path NTT run on every r-th row (same buffer, row pitch r times larger: api.cu).  Step 2 is this
// kernel: for every k2, read the r ADJACENT rows r*k2 + n1, multiply by the twiddles, do the order-r transform and write the r
// results to rows k2 + M*k1 of the destination (natural order; out of place).  HBM-bound: one read and one write of the array.
// The twiddle w_N^e is split by the CRT into (a root of order r) * (a root of order M): e = n1*k2,
//     w_N^e = GF_Root(r)^(a*e mod r) * GF_Root(M)^(b*e mod M),     a = M^-1 mod r,  b = r^-1 mod M,
// so it comes from a 9-entry constant table and the 2^20-entry power table the other kernels use.
#include "mixed_radix.h"
#include "gf.cuh"

namespace fecc {

struct RadixConsts {
    uint32_t zeta[9][3];       // GF_Root(r)^j (direction applied): {w, Whi, Wlo}
    uint32_t c1[3], c2[3];     // order-3 codelet: (X + X^2)/2, (X - X^2)/2 with X = GF_Root(3) (direction applied)
};

__device__ __forceinline__ uint32_t mulc(uint32_t x, const uint32_t (&t)[3]) { return gf::canon(gf::mul(x, t[0], t[1], t[2])); }
__device__ __forceinline__ uint32_t addm(uint32_t a, uint32_t b) { const uint32_t s = a + b; return (s < a || s >= gf::P) ? s - gf::P : s; }
__device__ __forceinline__ uint32_t subm(uint32_t a, uint32_t b) { return a >= b ? a - b : a + (gf::P - b); }

// order-3 transform of canonical residues (the arithmetic of the reference's NTT3 codelet, ntt.cpp:26-46)
__device__ __forceinline__ void dft3(uint32_t& f0, uint32_t& f1, uint32_t& f2, const RadixConsts& C)
{
    const uint32_t s = addm(f1, f2), d = subm(f1, f2);
    const uint32_t u = addm(f0, mulc(s, C.c1)), v = mulc(d, C.c2);
    f0 = addm(f0, s); f1 = addm(u, v); f2 = subm(u, v);
}

template <int R>
__global__ void __launch_bounds__(256) radix_pass_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint32_t pitch4, uint32_t s4, uint32_t M, uint32_t log_m,
                                                         uint32_t a_inv, uint32_t b_inv, int inverse, RadixConsts C, const uint4* __restrict__ tw)
{
    const size_t total = (size_t)M * s4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t k2 = (uint32_t)(i / s4), c = (uint32_t)(i - (size_t)k2 * s4);
        uint32_t f[R][4];
#pragma unroll
        for (int n1 = 0; n1 < R; ++n1) {
            const uint4 v = src[((size_t)R * k2 + n1) * pitch4 + c];
            f[n1][0] = v.x; f[n1][1] = v.y; f[n1][2] = v.z; f[n1][3] = v.w;
            const unsigned long long e = (unsigned long long)n1 * k2;
            const uint32_t er = (uint32_t)((a_inv * (e % R)) % R); uint32_t em = (uint32_t)(((unsigned long long)b_inv * (e & (M - 1))) & (M - 1));
            if (inverse) em = (M - em) & (M - 1);                               // the power table is forward; C.zeta already has the direction
            const uint4 t = tw[(size_t)em << (20 - log_m)];                    // GF_Root(M)^(+-em)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t x = gf::canon(gf::mul(f[n1][j], t.x, t.y, t.z));      // inputs are taken mod P
                f[n1][j] = er ? mulc(x, C.zeta[er]) : x;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (R == 3) {
                dft3(f[0][j], f[1][j], f[2][j], C);
            } else {                                                            // 3 x 3 four-step, as the reference's NTT9 (ntt.cpp:114-146)
                dft3(f[0][j], f[3][j], f[6][j], C); dft3(f[1][j], f[4][j], f[7][j], C); dft3(f[2][j], f[5][j], f[8][j], C);
                f[4][j] = mulc(f[4][j], C.zeta[1]); f[5][j] = mulc(f[5][j], C.zeta[2]);
                f[7][j] = mulc(f[7][j], C.zeta[2]); f[8][j] = mulc(f[8][j], C.zeta[4]);
                dft3(f[0][j], f[1][j], f[2][j], C); dft3(f[3][j], f[4][j], f[5][j], C); dft3(f[6][j], f[7][j], f[8][j], C);
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
            const int p = R == 3 ? k1 : (k1 % 3) * 3 + k1 / 3;                  // the 3 x 3 transpose: X[k1] sits at position 3*(k1 % 3) + k1 / 3
            dst[((size_t)k2 + (size_t)M * k1) * pitch4 + c] = make_uint4(f[p][0], f[p][1], f[p][2], f[p][3]);
        }
    }
}

static void fill3(uint32_t (&t)[3], uint32_t w) { const gf::Tw x = gf::make_tw(w); t[0] = x.w; t[1] = x.whi; t[2] = x.wlo; }

cudaError_t launch_radix_pass(const uint32_t* src, uint32_t* dst, uint32_t pitch4, uint32_t s4, uint32_t r, uint32_t M, bool inverse, const uint4* tw, int num_sms, cudaStream_t st)
{
    if ((r != 3 && r != 9) || !M || (M & (M - 1)) || M > gf::M) return cudaErrorInvalidValue;
    uint32_t log_m = 0; while ((1u << log_m) < M) ++log_m;
    RadixConsts C;
    uint32_t zr = gf::powmod(gf::GEN, (gf::P - 1) / r);                        // GF_Root(r), GF(p).cpp:267-276
    uint32_t z3 = gf::powmod(gf::GEN, (gf::P - 1) / 3);
    if (inverse) { zr = gf::inv(zr); z3 = gf::inv(z3); }
    for (uint32_t j = 0; j < 9; ++j) fill3(C.zeta[j], gf::powmod(zr, j));
    const uint32_t inv2 = (gf::P + 1) / 2, x1 = z3, x2 = gf::mulmod(z3, z3);
    fill3(C.c1, gf::mulmod((uint32_t)(((uint64_t)x1 + x2) % gf::P), inv2));
    fill3(C.c2, gf::mulmod((uint32_t)(((uint64_t)x1 + gf::P - x2) % gf::P), inv2));
    uint32_t a_inv = 0, b_inv = 0;                                              // a = M^-1 mod r, b = r^-1 mod M
    for (uint32_t a = 0; a < r; ++a) if (((uint64_t)a * (M % r)) % r == 1 % r) { a_inv = a; break; }
    if (M > 1) { uint64_t b = 1; for (uint32_t i = 0; i < log_m; ++i) b = (b * (2 - (r * b) % M + M)) % M; b_inv = (uint32_t)b; }     // Newton: r*b = 1 mod 2^k
    const size_t total = (size_t)M * s4;
    const unsigned grid = (unsigned)((total + 255) / 256 < (size_t)num_sms * 16 ? (total + 255) / 256 : (size_t)num_sms * 16);
    if (r == 3) radix_pass_kernel<3><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), pitch4, s4, M, log_m, a_inv, b_inv, inverse ? 1 : 0, C, tw);
    else        radix_pass_kernel<9><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), pitch4, s4, M, log_m, a_inv, b_inv, inverse ? 1 : 0, C, tw);
    return cudaGetLastError();
}

} // namespace fecc
