// The sm_100a pass kernel: one launch = one pass of the four-step / MFA decomposition (MFA_NTT, ntt.cpp:382-447)
// over the whole [N][SIZE] array, or -- with two transforms fused on a tile -- the end of the inverse transform,
// the RS scaling (RS.cpp:51-59) and the start of the forward transform in a single pass.
//
// Execution model: persistent CTAs of 256 threads, two per SM, each looping over work items
// (a row set x a few consecutive 64 KiB column strips).  Per tile: 16-byte cp.async (LDGSTS) loads straight into the
// bank-conflict-free shared-memory layout, the per-set twiddle table gathered from the L2-resident power table
// while those loads are in flight, ceil(LR/4) register rounds separated by one barrier each, and 128-bit stores
// from registers.  The kernel is bound by the integer multiply pipe (IMAD.HI), not by HBM: see profiles/.
#include "ntt_tile.cuh"
#include "ntt_pass.h"
#include <cuda_runtime.h>

namespace fecc {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 2) ntt_pass_kernel(const PassParams P)
{
    extern __shared__ __align__(128) uint4 smem[];
    uint4* tile = smem;                                   // 4096 chunks
    uint4* tw0  = smem + kTileChunks;                     // R entries
    uint4* tw1  = tw0 + (1u << P.log_r);                  // R entries (only when nxf == 2)

    const uint32_t tid  = threadIdx.x;
    const uint32_t zero = gf::opaque_zero();
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const uint32_t nitems = P.nsets * groups;
    const uint32_t nrounds = num_rounds(P.log_r);
    const uint4* src4 = reinterpret_cast<const uint4*>(P.src);

    for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const uint32_t set = item / groups, sg = item - set * groups;
        const uint32_t strip0 = sg * P.strips_per_item;
        const uint32_t strip1 = min(strip0 + P.strips_per_item, P.nstrips);
        for (uint32_t strip = strip0; strip < strip1; ++strip) {
            // (1) asynchronous tile load
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                unsigned long long sc; uint32_t ti;
                if (load_map(P, set, strip, tid, m, sc, ti)) cp_async16(tile + ti, src4 + sc);
            }
            // (2) twiddle tables of this row set, gathered from L2 while the tile is in flight
            if (strip == strip0) {
                build_table(P, 0, set, tid, tw0);
                if (P.nxf == 2) build_table(P, 1, set, tid, tw1);
            }
            cp_async_wait_all();
            __syncthreads();
            // (3) register rounds
            for (uint32_t xfi = 0; xfi < P.nxf; ++xfi) {
                const uint4* tws = xfi ? tw1 : tw0;
                for (uint32_t k = 0; k < nrounds; ++k) {
                    run_round(P, xfi, k, tid, set, strip, tile, tws, zero);
                    __syncthreads();
                }
            }
        }
    }
}

size_t pass_smem_bytes(const PassParams& P)
{
    return (size_t)kTileBytes + (size_t)P.nxf * ((size_t)16 << P.log_r);
}

cudaError_t launch_pass(const PassParams& P, int num_sms, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes + 2 * (16 << kMaxLogR));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const unsigned long long nitems = (unsigned long long)P.nsets * groups;
    unsigned long long grid = (unsigned long long)num_sms * 2;
    if (grid > nitems) grid = nitems;
    if (grid == 0) return cudaSuccess;
    ntt_pass_kernel<<<(unsigned)grid, kThreads, pass_smem_bytes(P), stream>>>(P);
    return cudaGetLastError();
}

} // namespace fecc
