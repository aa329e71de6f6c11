// Where do the cycles of a round go?  Runs the production round body (ntt_tile.cuh) in isolation:
//   mode 0: butterflies only (registers + twiddle LDS), no tile traffic, no barriers
//   mode 1: + tile LDS/STS every round, no barrier
//   mode 2: + tile LDS/STS + __syncthreads (the real round structure, minus global traffic)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../fastecc_b200/csrc/ntt_tile.cuh"
using namespace fecc;

template <int MODE, int LRT>
__global__ void __launch_bounds__(256, 2) rb(PassParams Pin, int iters, uint32_t k, uint32_t* sink)
{
    PassParams P = Pin; P.log_r = LRT; P.nxf = 1; P.parity = (LRT == 10);
    extern __shared__ __align__(128) uint4 smem[];
    uint4* tile = smem; uint4* tw = smem + kTileChunks;
    const uint32_t tid = threadIdx.x, zero = gf::opaque_zero();
    for (uint32_t i = tid; i < kTileChunks; i += 256) tile[i] = make_uint4(i * 2654435761u, i ^ 0x1234567, i * 40503u, i + 99);
    for (uint32_t i = tid; i < (1u << P.log_r); i += 256) tw[i] = P.tw[(i * 977u) & (gf::M - 1)];
    __syncthreads();
    RoundRegs r;
    Step st; st.xfi = 0; st.k = k; st.fused = false;
    round_read(P, k, 0, tid, tile, r);
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) round_read(P, k, 0, tid, tile, r);
        round_math(P, st, tid, 1 /*set (non-plain via t1)*/, tw, tw, r, zero);
        if (MODE >= 1) round_write_tile(P, k, 0, tid, tile, r);
        if (MODE >= 2) __syncthreads();
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc ^= r.x[i].x ^ r.x[i].y;
    sink[blockIdx.x * 256 + tid] = acc;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    uint4* tw; cudaMalloc(&tw, 16 << 20);
    { uint32_t* h = (uint32_t*)malloc(16 << 20); for (size_t i = 0; i < (4u << 20); i++) h[i] = (uint32_t)(i * 2654435761u) | 1; cudaMemcpy(tw, h, 16 << 20, cudaMemcpyHostToDevice); }
    uint32_t* sink; cudaMalloc(&sink, sms * 2 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int smem = kTileBytes + (16 << 10) ;
    cudaFuncSetAttribute(rb<0,10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rb<1,10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rb<2,10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rb<0,9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rb<1,9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rb<2,9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (uint32_t LR : {10u, 9u}) {
        PassParams P{}; P.tw = tw; P.log_r = LR; P.nxf = 1; P.parity = (LR == 10); P.xf[0] = Xform{12345, 0, 777};
        P.s4 = 256; P.pitch4 = 256; P.nstrips = 1;
        for (uint32_t k : {0u, 1u}) for (int ctas : {1, 2}) {
            const int iters = 200, grid = sms * ctas;
            auto run = [&](int mode) {
                float best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    cudaEventRecord(e0);
                    if (LR == 10) {
                    if (mode == 0) rb<0,10><<<grid, 256, smem>>>(P, iters, k, sink);
                    if (mode == 1) rb<1,10><<<grid, 256, smem>>>(P, iters, k, sink);
                    if (mode == 2) rb<2,10><<<grid, 256, smem>>>(P, iters, k, sink);
                    } else {
                    if (mode == 0) rb<0,9><<<grid, 256, smem>>>(P, iters, k, sink);
                    if (mode == 1) rb<1,9><<<grid, 256, smem>>>(P, iters, k, sink);
                    if (mode == 2) rb<2,9><<<grid, 256, smem>>>(P, iters, k, sink);
                    }
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                }
                const double warp_bfly_per_smsp = (double)iters * ((k == 0 || LR == 10) ? 160 : 128) /*bfly per thread per round*/ * 8 * ctas / 4;
                printf("LR=%u round k=%u ctas/SM=%d mode=%d: %.3f ms  -> %.2f cycles per warp-butterfly per SMSP (@1.965GHz)  [%s]\n", LR, k, ctas, mode, best,
                       best * 1e-3 * 1.965e9 / warp_bfly_per_smsp, cudaGetErrorString(cudaGetLastError()));
            };
            run(0); run(1); run(2);
        }
    }
    return 0;
}
