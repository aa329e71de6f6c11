// Byte blocks <-> GF(P) words.  The encoder needs every 32-bit word < P = 0xFFF00001; arbitrary bytes
// are not.  The reference describes (does not implement) a bit-optimal recoding, GF.md:72-104: view the top 12 bits of
// the W <= 1024 words of a block as digits in base 4096 and rewrite the digit string in base 4095 (no digit 0xFFF, so
// every word <= 0xFFEFFFFF < P) at the cost of ONE extra bit, stored as word W of the block (4096 B -> 4100 B):
//   extra = 0: no digit was 0xFFF, the block is unchanged;
//   extra = 1: the string starts with k index entries, `position of the j-th 0xFFF | (more follow) << 10`, followed by
//              the W - k digits that are not 0xFFF, in order.  (k entries + W - k digits = W digits again; an index
//              entry is < 2^11, a valid base-4095 digit.)  The low 20 bits of every word stay where they are.
// One CTA per block, one uint4 (4 words) per thread; blocks without a 0xFFF digit (78 % of random 4 KiB blocks) take the
// copy path after one __syncthreads_or.  Streaming kernels: 4096 B read + 4100 B written per block, HBM-bound.
#include "byte_recode.h"

namespace fecc {

constexpr uint32_t kDigitMax = 0xFFFu;
constexpr int kRecodeThreads = 256;                        // W / 4 <= 256 of them hold words

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane)
{
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, d); if (lane >= (uint32_t)d) s += t; }
    return s - v;
}
// exclusive prefix sum over the CTA; total returned through `total`
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ex = warp_excl_scan(v, lane);
    if (lane == 31) warp_sums[warp] = ex + v;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kRecodeThreads / 32; ++w) { const uint32_t s = warp_sums[w]; if ((uint32_t)w < warp) base += s; tot += s; }
    __syncthreads();                                       // warp_sums may be reused by the caller
    total = tot;
    return base + ex;
}

__global__ void __launch_bounds__(kRecodeThreads) bytes_to_gfp_kernel(const uint4* __restrict__ src, uint32_t* __restrict__ dst,
                                                                      uint32_t W, size_t pitch_words)
{
    __shared__ uint32_t digits[1024];
    __shared__ uint32_t warp_sums[kRecodeThreads / 32];
    const size_t blk = blockIdx.x;
    const uint32_t t = threadIdx.x, q = W / 4;
    const bool has = t < q;
    uint4 w = has ? src[blk * q + t] : make_uint4(0, 0, 0, 0);
    uint32_t x[4] = {w.x, w.y, w.z, w.w};
    uint32_t cnt = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) cnt += (has && (x[c] >> 20) == kDigitMax) ? 1u : 0u;
    uint4* out = reinterpret_cast<uint4*>(dst + blk * pitch_words);
    if (!__syncthreads_or((int)cnt)) {                     // nothing to recode
        if (has) out[t] = w;
        if (t == 0) dst[blk * pitch_words + W] = 0u;
        return;
    }
    uint32_t k;
    const uint32_t before = block_excl_scan(cnt, warp_sums, k);       // 0xFFF digits in front of this thread's words
    uint32_t seen = before;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (!has) break;
        const uint32_t i = 4 * t + c, d = x[c] >> 20;
        if (d == kDigitMax) { digits[seen] = i | ((seen + 1 < k) ? 0x400u : 0u); ++seen; }
        else                digits[k + i - seen] = d;
    }
    __syncthreads();
    if (has) {
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = (digits[4 * t + c] << 20) | (x[c] & 0xFFFFFu);
        out[t] = make_uint4(x[0], x[1], x[2], x[3]);
    }
    if (t == 0) dst[blk * pitch_words + W] = 1u;
}

__global__ void __launch_bounds__(kRecodeThreads) gfp_to_bytes_kernel(const uint32_t* __restrict__ src, uint4* __restrict__ dst,
                                                                      uint32_t W, size_t pitch_words)
{
    __shared__ uint32_t digits[1024];
    __shared__ uint32_t mark[1024];
    __shared__ uint32_t warp_sums[kRecodeThreads / 32];
    __shared__ uint32_t k_sh;
    const size_t blk = blockIdx.x;
    const uint32_t t = threadIdx.x, q = W / 4;
    const bool has = t < q;
    const uint4* in = reinterpret_cast<const uint4*>(src + blk * pitch_words);
    uint4 w = has ? in[t] : make_uint4(0, 0, 0, 0);
    const uint32_t extra = src[blk * pitch_words + W];     // same word for the whole CTA
    if (extra == 0) { if (has) dst[blk * q + t] = w; return; }
    uint32_t x[4] = {w.x, w.y, w.z, w.w};
    if (t == 0) k_sh = W;                                  // corrupt input (no terminator): every entry is an index
#pragma unroll
    for (int c = 0; c < 4; ++c) { digits[4 * t + c] = has ? (x[c] >> 20) : 0u; mark[4 * t + c] = 0u; }
    __syncthreads();
    if (has) {                                             // k - 1 = first entry whose "more follow" flag is clear
#pragma unroll
        for (int c = 0; c < 4; ++c) if (!(digits[4 * t + c] & 0x400u)) atomicMin(&k_sh, 4 * t + c + 1);
    }
    __syncthreads();
    const uint32_t k = k_sh;
    if (has) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { const uint32_t j = 4 * t + c; if (j < k) { const uint32_t pos = digits[j] & 0x3FFu; if (pos < W) mark[pos] = 1u; } }
    }
    __syncthreads();
    uint32_t cnt = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) cnt += has ? mark[4 * t + c] : 0u;
    uint32_t total;
    uint32_t seen = block_excl_scan(cnt, warp_sums, total);
    if (has) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t i = 4 * t + c;
            uint32_t d;
            if (mark[i]) { d = kDigitMax; ++seen; }
            else         d = digits[(k + i - seen) & 1023u];
            x[c] = (d << 20) | (x[c] & 0xFFFFFu);
        }
        dst[blk * q + t] = make_uint4(x[0], x[1], x[2], x[3]);
    }
}

cudaError_t launch_bytes_to_gfp(const void* d_bytes, uint32_t* d_words, size_t n_blocks, uint32_t W, size_t pitch_words, cudaStream_t stream)
{
    if (n_blocks == 0) return cudaSuccess;
    bytes_to_gfp_kernel<<<(unsigned)n_blocks, kRecodeThreads, 0, stream>>>(static_cast<const uint4*>(d_bytes), d_words, W, pitch_words);
    return cudaGetLastError();
}
cudaError_t launch_gfp_to_bytes(const uint32_t* d_words, void* d_bytes, size_t n_blocks, uint32_t W, size_t pitch_words, cudaStream_t stream)
{
    if (n_blocks == 0) return cudaSuccess;
    gfp_to_bytes_kernel<<<(unsigned)n_blocks, kRecodeThreads, 0, stream>>>(d_words, static_cast<uint4*>(d_bytes), W, pitch_words);
    return cudaGetLastError();
}

} // namespace fecc
