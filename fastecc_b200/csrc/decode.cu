// Erasure decoding behind the C ABI (include/fastecc_b200.h: fastecc_b200_rs_decode_*), SURVEY 8f rank 4.
//
// The reference describes the algorithm and does not implement it (README.md:88-119 "Fastest", RS.md:42-79, roadmap
// README.md:173).  Code word c[m] = f(rho^m), m < 2N, rho = GF_Root(2N), deg f < N; the encoder's layout is c[2i] = data
// block i, c[2j+1] = parity block j (RS.cpp:22-68).  For an erased set E (|E| <= N):
//     l(x) = prod_{e in E} (x - rho^e),   p = f*l is known at all 2N points (0 on E),   f(rho^e) = p'(rho^e) / l'(rho^e).
// Per word column: scale row m by l(rho^m), inverse NTT(2N), scale row t by t, forward NTT(2N), scale the erased rows by
// 1 / (2N * D[e]) with D = NTT(t * l_t).  Everything that depends only on WHICH rows are lost is built once
// (fastecc_b200_rs_decode_pattern) entirely on the device: the locator is a product tree over the erased points --
// leaves of 32 points by direct multiplication in one kernel, every further level ONE batched forward transform of all the
// polynomials (as word columns), one pairwise product (with the 1/n of the inverse transform folded in) and ONE batched
// inverse transform.  The Python mirror of the same algorithm is fastecc_b200/decoder.py.
#include "../../include/fastecc_b200.h"
#include "gf.cuh"
#include <cuda_runtime.h>
#include <vector>
#include <new>

namespace fecc {
const uint4* context_power_table();          // api.cu: g^e, e in [0, 2^20), or nullptr before fastecc_b200_init()
int          context_num_sms();
int          api_fail(int code, const char* fmt, ...);
}

struct fastecc_b200_erasures {
    size_t n2 = 0, me = 0;
    uint32_t* d_idx = nullptr;       // [me]  erased positions, ascending
    uint32_t* d_lv = nullptr;        // [n2]  l(rho^m): zero exactly on the erased rows
    uint32_t* d_ce = nullptr;        // [me]  1 / (2N * D[e])
    uint32_t* d_pos = nullptr;       // [n2]  0, 1, 2, ...: the constants of the formal derivative
};

namespace {

using namespace fecc;
constexpr uint32_t P = gf::P;
constexpr uint32_t kLeaf = 32;       // points per leaf polynomial

__device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
__device__ __forceinline__ uint32_t submod(uint32_t a, uint32_t b) { return a >= b ? a - b : a + (P - b); }
__device__ uint32_t invmod(uint32_t x)
{
    uint32_t r = 1;
    for (uint32_t e = P - 2; e; e >>= 1) { if (e & 1u) r = mulmod(r, x); x = mulmod(x, x); }
    return r;
}

#define GRID(n) (unsigned)(((n) + 255) / 256 < 8192 ? ((n) + 255) / 256 : 8192)
#define FOR_EACH(i, n) for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__global__ void iota_kernel(uint32_t* out, size_t n) { FOR_EACH(i, n) out[i] = (uint32_t)i; }

// Leaves: polynomial g (column g of `polys`, coefficient k in row k, pitch words per row) = prod over its <= 32 points of
// (x - rho^e); groups beyond the erased set are the constant 1.  rho^e = g^(e * 2^20 / n2) comes from the power table.
__global__ void leaves_kernel(const uint32_t* __restrict__ idx, size_t me, uint32_t log_step, const uint4* __restrict__ tw,
                              uint32_t* __restrict__ polys, size_t groups, size_t pitch)
{
    FOR_EACH(g, groups) {
        uint32_t c[kLeaf + 1];
#pragma unroll
        for (uint32_t k = 0; k <= kLeaf; ++k) c[k] = k == 0 ? 1u : 0u;
        uint32_t deg = 0;
        for (uint32_t t = 0; t < kLeaf; ++t) {
            const size_t e = g * kLeaf + t;
            if (e >= me) break;
            const uint32_t r = tw[((size_t)idx[e] << log_step) & (gf::M - 1)].x;      // rho^e
            ++deg;
#pragma unroll
            for (int k = kLeaf; k >= 1; --k) if ((uint32_t)k <= deg) c[k] = submod(c[k - 1], mulmod(c[k], r));       // c(x) * (x - r)
            c[0] = submod(0u, mulmod(c[0], r));
        }
#pragma unroll
        for (uint32_t k = 0; k <= kLeaf; ++k) polys[k * pitch + g] = c[k];
    }
}

// b[r][c] = a[r][2c] * a[r][2c+1] * scale   (point-wise product of neighbouring polynomials in the transform domain)
__global__ void pairmul_kernel(const uint32_t* __restrict__ a, size_t a_pitch, uint32_t* __restrict__ b, size_t b_pitch, size_t rows, size_t cols_out, uint32_t scale)
{
    FOR_EACH(i, rows * cols_out) {
        const size_t r = i / cols_out, c = i - r * cols_out;
        b[r * b_pitch + c] = mulmod(mulmod(a[r * a_pitch + 2 * c], a[r * a_pitch + 2 * c + 1]), scale);
    }
}
// dst (rows_dst x cols, zero padded) <- rows [0, rows_src) of src
__global__ void pad_rows_kernel(const uint32_t* __restrict__ src, size_t src_pitch, size_t rows_src, uint32_t* __restrict__ dst, size_t dst_pitch, size_t rows_dst, size_t cols)
{
    FOR_EACH(i, rows_dst * dst_pitch) {
        const size_t r = i / dst_pitch, c = i - r * dst_pitch;
        dst[i] = (r < rows_src && c < cols) ? src[r * src_pitch + c] : 0u;
    }
}
// column 0 of a pitched array -> two pitched single-column arrays: l_t and t * l_t
__global__ void locator_columns_kernel(const uint32_t* __restrict__ polys, size_t pitch, size_t ncoef, uint32_t* __restrict__ lv, uint32_t* __restrict__ dl, size_t n2)
{
    FOR_EACH(t, n2) {
        const uint32_t c = t < ncoef ? polys[t * pitch] : 0u;
        lv[t * 4] = c; lv[t * 4 + 1] = 0; lv[t * 4 + 2] = 0; lv[t * 4 + 3] = 0;
        dl[t * 4] = mulmod(c, (uint32_t)(t % P)); dl[t * 4 + 1] = 0; dl[t * 4 + 2] = 0; dl[t * 4 + 3] = 0;
    }
}
__global__ void finish_pattern_kernel(const uint32_t* __restrict__ lv4, const uint32_t* __restrict__ dl4, const uint32_t* __restrict__ idx, size_t n2, size_t me,
                                      uint32_t* __restrict__ lv, uint32_t* __restrict__ ce)
{
    FOR_EACH(i, n2) {
        lv[i] = lv4[i * 4];
        if (i < me) ce[i] = invmod(mulmod(dl4[(size_t)idx[i] * 4], (uint32_t)(n2 % P)));       // 1 / (2N * D[e])
    }
}
// out row i = code row idx[i] * ce[i]  (16-byte chunks)
__global__ void gather_scale_kernel(const uint4* __restrict__ code, size_t pitch4, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ ce,
                                    uint4* __restrict__ out, size_t out_pitch4, size_t me, uint32_t s4)
{
    for (size_t row = blockIdx.x; row < me; row += gridDim.x) {
        const uint32_t c = ce[row];
        const uint4* src = code + (size_t)idx[row] * pitch4;
        uint4* dst = out + row * out_pitch4;
        for (uint32_t k = threadIdx.x; k < s4; k += blockDim.x) {
            uint4 v = src[k];
            v.x = mulmod(v.x % P, c); v.y = mulmod(v.y % P, c); v.z = mulmod(v.z % P, c); v.w = mulmod(v.w % P, c);
            dst[k] = v;
        }
    }
}

struct Scratch {                      // frees what a failed / finished pattern build allocated
    std::vector<void*> p;
    ~Scratch() { for (void* q : p) cudaFree(q); }
    uint32_t* words(size_t n) { void* q = nullptr; if (cudaMalloc(&q, (n ? n : 1) * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return nullptr; } p.push_back(q); return (uint32_t*)q; }
};
inline size_t pad4(size_t n) { return (n + 3) / 4 * 4; }

#define TRY_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return api_fail(FASTECC_B200_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } while (0)
#define TRY_RC(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

int build_pattern(fastecc_b200_erasures* E, const uint32_t* erased, cudaStream_t st)
{
    const size_t n2 = E->n2, me = E->me;
    const uint4* tw = context_power_table();
    uint32_t log_n2 = 0; while (((size_t)1 << log_n2) < n2) ++log_n2;
    TRY_CUDA(cudaMalloc((void**)&E->d_pos, n2 * 4));
    iota_kernel<<<GRID(n2), 256, 0, st>>>(E->d_pos, n2);
    if (me == 0) return 0;
    TRY_CUDA(cudaMalloc((void**)&E->d_idx, me * 4));
    TRY_CUDA(cudaMalloc((void**)&E->d_lv, n2 * 4));
    TRY_CUDA(cudaMalloc((void**)&E->d_ce, me * 4));
    TRY_CUDA(cudaMemcpyAsync(E->d_idx, erased, me * 4, cudaMemcpyHostToDevice, st));
    Scratch tmp;
    // leaves
    size_t groups = 1; while (groups * kLeaf < me) groups *= 2;
    size_t d = kLeaf, cnt = groups, pitch = pad4(cnt);
    uint32_t* polys = tmp.words((d + 1) * pitch);
    if (!polys) return api_fail(FASTECC_B200_ENOMEM, "fastecc_b200_rs_decode_pattern: out of device memory");
    leaves_kernel<<<GRID(groups), 256, 0, st>>>(E->d_idx, me, 20 - log_n2, tw, polys, groups, pitch);
    // product tree: cnt polynomials of degree <= d  ->  cnt/2 of degree <= 2d
    while (cnt > 1) {
        const size_t rows = 4 * d, half = cnt / 2, hp = pad4(half);
        uint32_t* a = tmp.words(rows * pitch);
        uint32_t* b = tmp.words(rows * hp);
        if (!a || !b) return api_fail(FASTECC_B200_ENOMEM, "fastecc_b200_rs_decode_pattern: out of device memory");
        pad_rows_kernel<<<GRID(rows * pitch), 256, 0, st>>>(polys, pitch, d + 1, a, pitch, rows, cnt);
        TRY_RC(fastecc_b200_ntt_u32_dev(a, rows, cnt, pitch, 0, st));
        if (hp != half) TRY_CUDA(cudaMemsetAsync(b, 0, rows * hp * sizeof(uint32_t), st));      // pad columns: transformed along, never used
        pairmul_kernel<<<GRID(rows * half), 256, 0, st>>>(a, pitch, b, hp, rows, half, gf::inv((uint32_t)rows));
        TRY_RC(fastecc_b200_ntt_u32_dev(b, rows, half, hp, 1, st));
        polys = b; pitch = hp; cnt = half; d *= 2;                             // rows [0, d] of b hold the products
    }
    // l(rho^m) and D = NTT(t * l_t): two single-column transforms of order n2 (16-byte rows)
    uint32_t* lv4 = tmp.words(n2 * 4);
    uint32_t* dl4 = tmp.words(n2 * 4);
    if (!lv4 || !dl4) return api_fail(FASTECC_B200_ENOMEM, "fastecc_b200_rs_decode_pattern: out of device memory");
    locator_columns_kernel<<<GRID(n2), 256, 0, st>>>(polys, pitch, d + 1 < n2 ? d + 1 : n2, lv4, dl4, n2);
    TRY_RC(fastecc_b200_ntt_u32_dev(lv4, n2, 1, 4, 0, st));
    TRY_RC(fastecc_b200_ntt_u32_dev(dl4, n2, 1, 4, 0, st));
    finish_pattern_kernel<<<GRID(n2), 256, 0, st>>>(lv4, dl4, E->d_idx, n2, me, E->d_lv, E->d_ce);
    TRY_CUDA(cudaGetLastError());
    TRY_CUDA(cudaStreamSynchronize(st));                                        // the scratch buffers are freed on return
    return 0;
}

} // namespace

extern "C" {

int fastecc_b200_rs_decode_pattern(size_t n_rows, const uint32_t* erased, size_t n_erased, void* stream, fastecc_b200_erasures** out)
{
    const char* who = "fastecc_b200_rs_decode_pattern";
    if (!out) return api_fail(FASTECC_B200_EINVAL, "%s: null output", who);
    *out = nullptr;
    if (!fecc::context_power_table()) return api_fail(FASTECC_B200_ENOINIT, "%s: call fastecc_b200_init() first", who);
    if (n_rows < 2 || (n_rows & (n_rows - 1)) || n_rows > ((size_t)1 << FASTECC_B200_MAX_LOG_N))
        return api_fail(FASTECC_B200_EINVAL, "%s: the code word must have 2N = 2 .. 2^20 rows (a power of two), got %zu", who, n_rows);
    if (n_erased > n_rows / 2) return api_fail(FASTECC_B200_EINVAL, "%s: at most N = %zu of the 2N rows can be erased, got %zu", who, n_rows / 2, n_erased);
    if (n_erased && !erased) return api_fail(FASTECC_B200_EINVAL, "%s: null position list", who);
    for (size_t i = 0; i < n_erased; ++i)
        if (erased[i] >= n_rows || (i && erased[i] <= erased[i - 1])) return api_fail(FASTECC_B200_EINVAL, "%s: erased positions must be ascending, distinct and below %zu", who, n_rows);
    fastecc_b200_erasures* E = new (std::nothrow) fastecc_b200_erasures;
    if (!E) return api_fail(FASTECC_B200_ENOMEM, "%s: out of host memory", who);
    E->n2 = n_rows; E->me = n_erased;
    const int rc = build_pattern(E, erased, (cudaStream_t)stream);
    if (rc) { fastecc_b200_rs_decode_free(E); return rc; }
    *out = E;
    return 0;
}

int fastecc_b200_rs_decode_recover(const fastecc_b200_erasures* E, uint32_t* d_code, size_t size, size_t pitch, uint32_t* d_out, size_t out_pitch, void* stream)
{
    const char* who = "fastecc_b200_rs_decode_recover";
    if (!E || !d_code) return api_fail(FASTECC_B200_EINVAL, "%s: null argument", who);
    if (E->me == 0) return 0;
    if (!d_out) return api_fail(FASTECC_B200_EINVAL, "%s: null output buffer", who);
    if (size == 0 || pitch < size || pitch % 4 || out_pitch < size || out_pitch % 4 || ((uintptr_t)d_code) % 16 || ((uintptr_t)d_out) % 16)
        return api_fail(FASTECC_B200_EINVAL, "%s: needs SIZE >= 1, 16-byte aligned buffers, pitches >= SIZE and multiples of 4 words", who);
    cudaStream_t st = (cudaStream_t)stream;
    TRY_RC(fastecc_b200_row_scale_dev(d_code, E->n2, size, pitch, E->d_lv, st));            // p(rho^m) = c[m] * l(rho^m): zero on E whatever was there
    TRY_RC(fastecc_b200_ntt_u32_dev(d_code, E->n2, size, pitch, 1, st));                     // 2N * coefficients of p
    TRY_RC(fastecc_b200_row_scale_dev(d_code, E->n2, size, pitch, E->d_pos, st));            // t * p_t: x * p'(x)
    TRY_RC(fastecc_b200_ntt_u32_dev(d_code, E->n2, size, pitch, 0, st));                     // 2N * rho^j * p'(rho^j)
    const unsigned grid = (unsigned)(E->me < (size_t)fecc::context_num_sms() * 8 ? E->me : (size_t)fecc::context_num_sms() * 8);
    gather_scale_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_code), pitch / 4, E->d_idx, E->d_ce, reinterpret_cast<uint4*>(d_out), out_pitch / 4,
                                              E->me, (uint32_t)((size + 3) / 4));
    TRY_CUDA(cudaGetLastError());
    return 0;
}

size_t fastecc_b200_rs_decode_count(const fastecc_b200_erasures* E) { return E ? E->me : 0; }

void fastecc_b200_rs_decode_free(fastecc_b200_erasures* E)
{
    if (!E) return;
    cudaFree(E->d_idx); cudaFree(E->d_lv); cudaFree(E->d_ce); cudaFree(E->d_pos);
    delete E;
}

} // extern "C"
