// The sm_100a pass kernel: one launch = one pass of the four-step / MFA decomposition (MFA_NTT, ntt.cpp:382-447)
// over the whole [N][SIZE] array, or -- with two transforms fused on a tile -- the end of the inverse transform,
// the RS scaling (RS.cpp:51-59) and the start of the forward transform in a single pass.
//
// Execution model: persistent CTAs of 256 threads, two per SM, each walking a sequence of
// 64 KiB tiles (row set x column strip).  Per tile:
//   * the tile and, when the row set changes, its twiddle table arrive by 16-byte cp.async (LDGSTS) straight into
//     the bank-conflict-free shared-memory layout -- issued one tile AHEAD, right after the previous tile's last
//     round has pulled its slots into registers, so the loads fly under a full four-stage round of butterflies;
//   * ceil(LR/5) <= 2 register rounds of up to five radix-2 stages (three for a fused two-transform tile, whose
//     middle round runs 4+5 stages back to back in registers) separated by one block barrier each;
//   * 128-bit stores of the finished rows straight from registers.
// The kernel is bound by the integer multiply pipe (IMAD.HI on "fmaheavy"), not by HBM: see profiles/.
#include "ntt_tile.cuh"
#include "ntt_pass.h"
#include <cuda_runtime.h>

namespace fecc {

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

struct TileIter {                       // the CTA's walk over (item -> strips)
    uint32_t item, set, strip, strip_end;
};
__device__ __forceinline__ bool iter_decode(const PassParams& P, uint32_t groups, uint32_t nitems, TileIter& it)
{
    if (it.item >= nitems) return false;
    it.set = it.item / groups;
    const uint32_t sg = it.item - it.set * groups;
    it.strip = sg * P.strips_per_item;
    it.strip_end = min(it.strip + P.strips_per_item, P.nstrips);
    return true;
}
__device__ __forceinline__ bool iter_next(const PassParams& P, uint32_t groups, uint32_t nitems, TileIter& it)
{
    if (++it.strip < it.strip_end) return true;
    it.item += gridDim.x;
    return iter_decode(P, groups, nitems, it);
}

// LR / NXF / PARITY are compile-time: the kernel patches them into its copy of the parameters, so every shift,
// stride and placement branch in ntt_tile.cuh folds to an immediate (keeps the 64 data registers from spilling).
template <int LR, int NXF, int PARITY>
__global__ void __launch_bounds__(kThreads, 2) ntt_pass_kernel(const PassParams Pin)
{
    PassParams P = Pin;
    P.log_r = LR; P.nxf = NXF; P.parity = PARITY;
    extern __shared__ __align__(128) uint4 smem[];
    uint4* tile = smem;                                   // 4096 chunks
    uint4* tabs = smem + kTileChunks;                     // [2 buffers][nxf][R]
    const uint32_t R = 1u << P.log_r;
    const uint32_t tid  = threadIdx.x;
    const uint32_t zero = gf::opaque_zero();
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const uint32_t nitems = P.nsets * groups;
    const uint32_t nsteps = num_steps(P.log_r, P.nxf);

    TileIter cur; cur.item = blockIdx.x;
    if (!iter_decode(P, groups, nitems, cur)) return;
    uint32_t tb = 0;                                      // table buffer used by the current tile

    load_tile(P, cur.set, cur.strip, tid, tile);          // prologue: first tile + its tables
    for (uint32_t x = 0; x < P.nxf; ++x) build_table(P, x, cur.set, tid, tabs + (tb * P.nxf + x) * R);

    for (;;) {
        cp_async_wait_all();
        __syncthreads();                                  // tile + tables of `cur` have landed
        const bool active = thread_active(P, tid, cur.strip);
        const uint4* tw0 = tabs + (tb * P.nxf) * R;
        const uint4* tw1 = tw0 + R;
        RoundRegs r;
        TileIter nxt = cur;
        bool has_next = false;
        for (uint32_t s = 0; s < nsteps; ++s) {
            const Step st = step_of(P.log_r, P.nxf, s);
            const bool last = s + 1 == nsteps;
            if (active) round_read(P, st.k, st.xfi, tid, tile, r);
            if (last) {
                __syncthreads();                          // every slot is in registers: the tile buffer is free
                has_next = iter_next(P, groups, nitems, nxt);
                if (has_next) {                           // prefetch under the last step's butterflies
                    load_tile(P, nxt.set, nxt.strip, tid, tile);
                    if (nxt.set != cur.set)
                        for (uint32_t x = 0; x < P.nxf; ++x) build_table(P, x, nxt.set, tid, tabs + ((tb ^ 1u) * P.nxf + x) * R);
                }
            }
            if (active) round_math(P, st, tid, cur.set, tw0, tw1, r, zero);
            if (!last) {
                if (active) round_write_tile(P, st.k, st.xfi, tid, tile, r);
                __syncthreads();
            } else if (active) {
                round_write_global(P, st, tid, cur.set, cur.strip, r);
            }
        }
        if (!has_next) break;
        if (nxt.set != cur.set) tb ^= 1u;
        cur = nxt;
    }
}

size_t pass_smem_bytes(const PassParams& P)
{
    return (size_t)kTileBytes + (size_t)2 * P.nxf * ((size_t)16 << P.log_r);
}

template <int LR, int NXF, int PARITY>
static cudaError_t launch_inst(const PassParams& P, unsigned grid, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<LR, NXF, PARITY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kTileBytes + 2 * NXF * (16 << LR));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    ntt_pass_kernel<LR, NXF, PARITY><<<grid, kThreads, pass_smem_bytes(P), stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_pass(const PassParams& P, int num_sms, cudaStream_t stream)
{
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const unsigned long long nitems = (unsigned long long)P.nsets * groups;
    unsigned long long grid = (unsigned long long)num_sms * 2;
    if (grid > nitems) grid = nitems;
    if (grid == 0) return cudaSuccess;
    const unsigned g = (unsigned)grid;
    if (P.parity != ((P.log_r == 10 && P.nxf == 1) ? 1u : 0u)) return cudaErrorInvalidValue;
#define FECC_CASE(L) case L: return P.nxf == 2 ? launch_inst<L, 2, 0>(P, g, stream) : launch_inst<L, 1, (L == 10)>(P, g, stream);
    switch (P.log_r) {
        FECC_CASE(5) FECC_CASE(6) FECC_CASE(7) FECC_CASE(8) FECC_CASE(9) FECC_CASE(10)
        default: return cudaErrorInvalidValue;
    }
#undef FECC_CASE
}

} // namespace fecc
