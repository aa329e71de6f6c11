// Where do the cycles of a round go?  Runs the production round body (ntt_tile.cuh) in isolation:
//   mode 0: butterflies only (registers + twiddle LDS), no tile traffic, no barriers
//   mode 1: + tile LDS/STS every round, no barrier
//   mode 2: + tile LDS/STS + __syncthreads (the real round structure, minus global traffic)
// Build variants: default (FP64 quotient, I2F), -DFECC_CVT_MAGIC, -DFECC_MUL_BARRETT.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../fastecc_b200/csrc/ntt_tile.cuh"
using namespace fecc;

template <int MODE, int LRT>
__global__ void __launch_bounds__(256, 2) rb(PassParams Pin, int iters, uint32_t k, uint32_t* sink)
{
    PassParams P = Pin; P.log_r = LRT; P.nxf = 1;
    extern __shared__ __align__(128) uint4 smem[];
    uint4* tile = smem; uint4* tw = smem + kTileChunks;
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < kTileChunks; i += 256) tile[i] = make_uint4(i * 2654435761u, i ^ 0x1234567, i * 40503u, i + 99);
    for (uint32_t i = tid; i < (1u << P.log_r); i += 256) tw[i] = stage_entry(P.tw[(i * 977u) & (gf::M - 1)]);
    __syncthreads();
    RoundRegs r;
    Step st; st.xfi = 0; st.k = k; st.fused = false;
    round_read(P, k, 0, tid, tile, r);
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) round_read(P, k, 0, tid, tile, r);
        round_math(P, st, tid, 1 /*set (non-plain via t1)*/, tw, tw, r);
        if (MODE >= 1) round_write_tile(P, k, 0, tid, tile, r);
        if (MODE >= 2) __syncthreads();
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc ^= r.x[i].x ^ r.x[i].y;
    sink[blockIdx.x * 256 + tid] = acc;
}

template <int MODE, int LRT>
static float run1(PassParams P, int grid, int smem, int iters, uint32_t k, uint32_t* sink)
{
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(rb<MODE, LRT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        rb<MODE, LRT><<<grid, 256, smem>>>(P, iters, k, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    uint4* tw; cudaMalloc(&tw, 16 << 20);
    { uint32_t* h = (uint32_t*)malloc(16 << 20); for (size_t i = 0; i < (4u << 20); i++) h[i] = (uint32_t)(i * 2654435761u) | 1; cudaMemcpy(tw, h, 16 << 20, cudaMemcpyHostToDevice); }
    uint32_t* sink; cudaMalloc(&sink, sms * 2 * 256 * 4);
    const int smem = kTileBytes + (16 << 10);
#if defined(FECC_MUL_BARRETT)
    const char* variant = "barrett";
#elif defined(FECC_CVT_MAGIC)
    const char* variant = "fp64-quotient/magic";
#else
    const char* variant = "fp64-quotient/i2f";
#endif
    for (uint32_t LR : {10u, 9u}) {
        PassParams P{}; P.tw = tw; P.log_r = LR; P.nxf = 1; P.xf[0] = Xform{12345, 0, 777};
        P.s4 = 256; P.pitch4 = 256; P.nstrips = 1;
        for (uint32_t k : {0u, 1u}) for (int ctas : {1, 2}) {
            const int iters = 200, grid = sms * ctas;
            for (int mode = 0; mode < 3; ++mode) {
                float ms = 0;
                if (LR == 10) ms = mode == 0 ? run1<0, 10>(P, grid, smem, iters, k, sink) : mode == 1 ? run1<1, 10>(P, grid, smem, iters, k, sink) : run1<2, 10>(P, grid, smem, iters, k, sink);
                else          ms = mode == 0 ? run1<0, 9>(P, grid, smem, iters, k, sink) : mode == 1 ? run1<1, 9>(P, grid, smem, iters, k, sink) : run1<2, 9>(P, grid, smem, iters, k, sink);
                const double warp_bfly_per_smsp = (double)iters * ((k == 0 || LR == 10) ? 160 : 128) /*bfly per thread per round*/ * 8 * ctas / 4;
                printf("%-20s LR=%u round k=%u ctas/SM=%d mode=%d: %.3f ms  -> %.2f cycles per warp-butterfly per SMSP (@1.965GHz)  [%s]\n", variant, LR, k, ctas, mode, ms,
                       ms * 1e-3 * 1.965e9 / warp_bfly_per_smsp, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
