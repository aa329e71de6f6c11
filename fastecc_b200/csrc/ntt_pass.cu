// The sm_100a pass kernel: one launch = one pass of the four-step / MFA decomposition (MFA_NTT, ntt.cpp:382-447)
// over the whole [N][SIZE] array, or -- with two transforms fused on a tile -- the end of the inverse transform,
// the RS scaling (RS.cpp:51-59) and the start of the forward transform in a single pass.
//
// Execution model: persistent CTAs of 256 threads, two per SM, each walking a sequence of
// 64 KiB tiles (row set x column strip).  Per tile:
//   * one elected thread issues the TMA traffic: cp.async.bulk.tensor (3-D tensor map over [word][row-in-set][set],
//     boxes of <= 256 rows) for the tile and, when the row set changes, one cp.async.bulk for the set's precomputed
//     twiddle tables; completion is tracked by an mbarrier (expect_tx / complete_tx).  The next tile is requested
//     right after the current tile's last step has pulled its slots into registers, so it flies under a full
//     five-stage round of butterflies; out-of-range columns are zero-filled by the TMA unit;
//   * ceil(LR/5) <= 2 register rounds of up to five radix-2 stages (three for a fused two-transform tile, whose
//     middle round runs 4+5 stages back to back in registers) separated by one block barrier each;
//   * 64-bit stores of the finished rows straight from registers.
// The kernel is bound by the integer multiply pipe (IMAD.HI on "fmaheavy"), not by HBM: see profiles/.
#include "ntt_tile.cuh"
#include "ntt_pass.h"
#include "plan.h"
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace fecc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, uint32_t c0, uint32_t c1, uint32_t c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The CTA's t-th tile.  Work items (a row set x strips_per_item consecutive strips) are dealt round-robin to the
// CTAs; plan.h guarantees strips_per_item divides nstrips, so tile t of a CTA is strip (t % spi) of its (t / spi)-th item.
__device__ __forceinline__ bool tile_decode(const PassParams& P, uint32_t groups, uint32_t nitems, uint32_t t, uint32_t& set, uint32_t& strip)
{
    const uint32_t spi = P.strips_per_item;
    const uint32_t item = blockIdx.x + (t / spi) * gridDim.x;
    if (item >= nitems) return false;
    set = item / groups;
    strip = (item - set * groups) * spi + (t % spi);
    return true;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }

// LR / NXF / TMA are compile-time: the kernel patches them into its copy of the parameters, so every shift, stride
// and placement branch in ntt_tile.cuh folds to an immediate (keeps the 64 data registers from spilling).
template <int LR, int NXF, int TMA>
__global__ void __launch_bounds__(kThreads, 2) ntt_pass_kernel(const PassParams Pin, const __grid_constant__ CUtensorMap tmap)
{
    PassParams P = Pin;
    P.log_r = LR; P.nxf = NXF; P.use_tma = TMA != 0;
    if (TMA != 2) P.log_g = 0;                            // TMA == 2: the instantiation whose stores pick a destination GPU (sharded, §8)
    extern __shared__ __align__(1024) uint4 smem[];
    constexpr uint32_t R = 1u << LR;
    uint4* tile = smem;                                   // 4096 chunks, natural row order
    // stage tables: a transform whose twist does not depend on the row set (t1 == 0) has ONE table, loaded once;
    // the others are double-buffered per set.  Slot of transform x in buffer b: tabs + (b*NXF + x)*R; set-independent
    // tables always use b = 0 (so a fused BC tile needs 3 slots, a plain A tile 1).
    uint4* tabs = smem + kTileChunks;
    const bool var0 = P.xf[0].t1 != 0, var1 = NXF == 2 && P.xf[1].t1 != 0;
    const uint32_t nslots = var1 ? 4u : (var0 ? (uint32_t)NXF + 1u : (uint32_t)NXF);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tabs + nslots * R);
    const uint32_t tid  = threadIdx.x;
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const uint32_t nitems = P.nsets * groups;
    constexpr uint32_t nsteps = NXF == 2 ? (LR > kStages ? 3u : 1u) : (LR > kStages ? 2u : 1u);
    constexpr uint32_t kWt = 16384u >> LR;                // words per tile row
    constexpr uint32_t kRowsBox = R < 256u ? R : 256u;    // TMA box: kRowsBox rows x kWt words
    constexpr uint32_t kBoxes = R / kRowsBox;

    uint32_t cur_set, cur_strip;
    if (!tile_decode(P, groups, nitems, 0, cur_set, cur_strip)) return;
    uint32_t tb = 0;                                      // table buffer used by the current tile
    uint32_t phase = 0;
    uint64_t* rbar = bar + 1;                             // "all eight warps have pulled the last step's slots into registers"

    // request tile (set, strip) [+ the set's tables into buffer tbuf]
    auto request = [&](uint32_t set, uint32_t strip, bool with_tables, uint32_t tbuf) {
        if (TMA) {
            if (tid == 0) {
                fence_proxy_async();                      // order our earlier generic-proxy accesses to the tile before the async-proxy writes
                const bool first = with_tables && tbuf == 2u;         // prologue: also the set-independent tables
                const uint32_t ld0 = with_tables && (var0 || first), ld1 = NXF == 2 && with_tables && (var1 || first);
                mbar_expect_tx(bar, kTileBytes + (ld0 + ld1) * (R * 16u));
#pragma unroll
                for (uint32_t b = 0; b < kBoxes; ++b)
                    tma_load_3d(tile + b * (kRowsBox * (kWt / 4)), &tmap, bar, strip * kWt, b * kRowsBox, set);
                const uint4* tsrc = P.tables + (size_t)set * P.table_set_stride;
                const uint32_t tbv = first ? 0u : tbuf;
                if (ld0) bulk_load(tabs + ((var0 ? tbv : 0u) * NXF + 0) * R, tsrc, R * 16u, bar);
                if (ld1) bulk_load(tabs + ((var1 ? tbv : 0u) * NXF + 1) * R, tsrc + R, R * 16u, bar);
            }
        } else {
            load_tile_cpasync(P, set, strip, tid, tile);
            if (with_tables) {
                const bool first = tbuf == 2u;
                const uint32_t tbv = first ? 0u : tbuf;
                const uint4* tsrc = P.tables + (size_t)set * P.table_set_stride;
                if (var0 || first) for (uint32_t i = tid; i < R; i += kThreads) copy16(tabs + ((var0 ? tbv : 0u) * NXF + 0) * R + i, tsrc + i);
                if (NXF == 2 && (var1 || first)) for (uint32_t i = tid; i < R; i += kThreads) copy16(tabs + ((var1 ? tbv : 0u) * NXF + 1) * R + i, tsrc + R + i);
            }
        }
    };

    if (TMA) {
        if (tid == 0) { mbar_init(bar, 1); mbar_init(rbar, kThreads / 32); fence_mbar_init(); }
        __syncthreads();
    }
    request(cur_set, cur_strip, true, 2u);                // prologue: first tile + all its tables (buffer 0)

    for (uint32_t t = 0;; ++t) {
        if (TMA) mbar_wait(bar, phase);
        else     { cp_async_wait_all(); __syncthreads(); }
        const bool active = thread_active(P, tid, cur_strip);
        RoundRegs r;
#pragma unroll 1
        for (uint32_t s = 0; s < nsteps; ++s) {
            const Step st = step_of(LR, NXF, s);
            const bool last = s + 1 == nsteps;
            if (active) round_read(P, st.k, st.xfi == 0, tid, tile, r);
            if (last) {                                   // every slot is in registers: the tile buffer is free.
                if (TMA) {                                // each warp reports in; only the issuing thread waits for all of them
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(rbar);
                    if (tid == 0) {                       // the next tile flies under the last step's butterflies
                        mbar_wait(rbar, phase);
                        uint32_t ns, nst;
                        if (tile_decode(P, groups, nitems, t + 1, ns, nst)) request(ns, nst, ns != cur_set, tb ^ 1u);
                    }
                } else {
                    __syncthreads();
                    uint32_t ns, nst;
                    if (tile_decode(P, groups, nitems, t + 1, ns, nst)) request(ns, nst, ns != cur_set, tb ^ 1u);
                }
            }
            const uint4* tw0 = tabs + ((var0 ? tb : 0u) * NXF) * R;
            const uint4* tw1 = tabs + ((var1 ? tb : 0u) * NXF + 1) * R;
            #if !defined(FECC_SKIP_MATH)                                   // -DFECC_SKIP_MATH: copy-only build for measuring the memory floor of a pass
            if (active) round_math(P, st, tid, cur_set, tw0, tw1, r);
#endif
            if (!last) {
                if (active) round_write_tile(P, st.k, st.xfi == 0, tid, tile, r);
                __syncthreads();
            } else if (active) {
                round_write_global(P, st, tid, cur_set, cur_strip, r);
            }
        }
        uint32_t nxt_set, nxt_strip;                      // (re)decoded here, where registers are plentiful
        if (!tile_decode(P, groups, nitems, t + 1, nxt_set, nxt_strip)) break;
        phase ^= 1u;
        if (nxt_set != cur_set) tb ^= 1u;
        cur_set = nxt_set; cur_strip = nxt_strip;
    }
}


__device__ __forceinline__ void group_sync(uint32_t g) { asm volatile("bar.sync %0, %1;" :: "r"(g + 1u), "r"((uint32_t)kThreads) : "memory"); }

// "Dual" schedule for single-transform passes (A, D, A', B'): ONE 512-thread CTA per SM = two groups of 256 threads
// that each run the CTA-level tile schedule of ntt_pass_kernel on alternate tiles of the CTA's sequence, sharing a pool
// of THREE tile buffers.  Tile k lives in buffer k % 3; when the group working on tile k has lifted it into registers
// for its last round, the buffer is free and the group's thread 0 requests tile k+3 (the OTHER group's tile after
// next) into it.  A tile's load therefore has a whole tile period of the other group to land, instead of the half
// period a CTA of ntt_pass_kernel can offer with its single buffer: the strided passes are balanced between DRAM
// (0.85 ms copy-only) and butterflies (~0.9 ms), so hiding one behind the other is what they need.
// Stage tables: a set-independent table is loaded once; otherwise each group double-buffers its own (slot 2g + parity),
// requested one tile ahead by the group itself.
template <int LR, int SHARD>
__global__ void __launch_bounds__(2 * kThreads, 1) ntt_pass_dual_kernel(const PassParams Pin, const __grid_constant__ CUtensorMap tmap)
{
    PassParams P = Pin;
    P.log_r = LR; P.nxf = 1; P.use_tma = 1;
    if (!SHARD) P.log_g = 0;
    extern __shared__ __align__(1024) uint4 smem[];
    constexpr uint32_t R = 1u << LR;
    constexpr uint32_t kBufs = 3;
    const bool var0 = P.xf[0].t1 != 0;
    uint4* tabs = smem + kBufs * kTileChunks;                                   // var0 ? [2 groups][2][R] : [R]
    // full[buffer][use & 1]: the tile of the buffer's use-th fill has landed (parity (use >> 1) & 1).  Two barriers per
    // buffer because the consumer of fill u+1 is the OTHER group, which may arrive while fill u is still in flight:
    // on a single barrier its parity wait would be satisfied by fill u-1.
    uint64_t* full  = reinterpret_cast<uint64_t*>(tabs + (var0 ? 4u : 1u) * R);
    uint64_t* rbar  = full + 2 * kBufs;                                         // rbar[2]: the group's 8 warps are done reading their tile
    uint64_t* tfull = rbar + 2;                                                 // tfull[2]: the group's next table has landed ([0] also: the shared table)
    const uint32_t g = threadIdx.x / kThreads, tid = threadIdx.x % kThreads;
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const uint32_t nitems = P.nsets * groups;
    constexpr uint32_t nsteps = LR > kStages ? 2u : 1u;
    constexpr uint32_t kWt = 16384u >> LR;
    constexpr uint32_t kRowsBox = R < 256u ? R : 256u;
    constexpr uint32_t kBoxes = R / kRowsBox;

    if (threadIdx.x == 0) {
        for (uint32_t b = 0; b < 2 * kBufs; ++b) mbar_init(full + b, 1);
        mbar_init(rbar, kThreads / 32); mbar_init(rbar + 1, kThreads / 32);
        mbar_init(tfull, 1); mbar_init(tfull + 1, 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto request_tile = [&](uint32_t k, uint32_t buf, uint32_t use) {           // one thread
        uint32_t set, strip;
        if (!tile_decode(P, groups, nitems, k, set, strip)) return;
        fence_proxy_async();
        uint64_t* bar = full + 2 * buf + (use & 1u);
        mbar_expect_tx(bar, kTileBytes);
        uint4* dst = smem + buf * kTileChunks;
#pragma unroll
        for (uint32_t b = 0; b < kBoxes; ++b)
            tma_load_3d(dst + b * (kRowsBox * (kWt / 4)), &tmap, bar, strip * kWt, b * kRowsBox, set);
    };
    auto request_table = [&](uint32_t k, uint32_t slot, uint64_t* bar) {        // one thread
        uint32_t set, strip;
        if (!tile_decode(P, groups, nitems, k, set, strip)) return;
        fence_proxy_async();
        mbar_expect_tx(bar, R * 16u);
        bulk_load(tabs + slot * R, P.tables + (size_t)set * P.table_set_stride, R * 16u, bar);
    };

    if (threadIdx.x == 0) {
        if (!var0) request_table(0, 0, tfull);
        request_tile(0, 0, 0); request_tile(1, 1, 0); request_tile(2, 2, 0);
    }
    if (var0 && tid == 0) request_table(g, 2 * g, tfull + g);
    if (!var0) mbar_wait(tfull, 0);

    uint32_t buf = g, use = 0;                            // tile k = g + 2*j: buffer k % 3, its (k / 3)-th use
    for (uint32_t k = g, j = 0;; k += 2, ++j) {
        uint32_t cur_set, cur_strip;
        if (!tile_decode(P, groups, nitems, k, cur_set, cur_strip)) break;
        const uint32_t tslot = var0 ? 2 * g + (j & 1u) : 0u;
        if (var0) mbar_wait(tfull + g, j & 1u);
        mbar_wait(full + 2 * buf + (use & 1u), (use >> 1) & 1u);
        uint4* tile = smem + buf * kTileChunks;
        const uint4* tw0 = tabs + tslot * R;
        const bool active = thread_active(P, tid, cur_strip);
        RoundRegs r;
#pragma unroll 1
        for (uint32_t s = 0; s < nsteps; ++s) {
            const Step st = step_of(LR, 1, s);
            const bool last = s + 1 == nsteps;
            if (active) round_read(P, st.k, true, tid, tile, r);
            if (last) {
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(rbar + g);
                if (tid == 0) {
                    mbar_wait(rbar + g, j & 1u);
                    request_tile(k + 3, buf, use + 1);                          // the other group's tile after next
                    if (var0) request_table(k + 2, 2 * g + ((j + 1) & 1u), tfull + g);   // our own next table: its slot was last used by tile k-2
                }
            }
            #if !defined(FECC_SKIP_MATH)
            if (active) round_math(P, st, tid, cur_set, tw0, tw0, r);
#endif
            if (!last) {
                if (active) round_write_tile(P, st.k, true, tid, tile, r);
                group_sync(g);
            } else if (active) {
                round_write_global(P, st, tid, cur_set, cur_strip, r);
            }
        }
        buf += 2; if (buf >= kBufs) buf -= kBufs;
        use = (k + 2) / 3;
    }
}

// One thread per table entry: tables[set][xfi][idx] = g^exponent (entry 0 of every table is unused).
__global__ void build_tables_kernel(const PassParams P, uint4* out, uint32_t nsets_tab)
{
    const uint32_t R = 1u << P.log_r;
    const size_t total = (size_t)nsets_tab * P.nxf * R;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t idx = (uint32_t)(i & (R - 1));
        const uint32_t xfi = (uint32_t)((i >> P.log_r) % P.nxf);
        const uint32_t set = (uint32_t)((i >> P.log_r) / P.nxf);
        out[i] = idx ? stage_entry(P.tw[table_entry_exponent(P, xfi, set, idx)]) : make_uint4(0, 0, 0, 0);
    }
}

cudaError_t launch_build_tables(PassParams& P, uint4* out, cudaStream_t stream)
{
    const uint32_t ns = table_sets(P);
    P.tables = out;
    P.table_set_stride = ns > 1 ? (P.nxf << P.log_r) : 0u;
    build_tables_kernel<<<592, 256, 0, stream>>>(P, out, ns);
    return cudaGetLastError();
}

// shared memory of the CTA-level kernel: tile + table slots + two mbarriers.  Table slot of transform x in buffer b is
// (b*nxf + x); buffer 1 is needed only up to the last set-dependent transform.
size_t pass_smem_bytes(const PassParams& P)
{
    uint32_t slots = P.nxf;                                            // buffer 0
    if (P.nxf == 2 && P.xf[1].t1) slots = 4; else if (P.xf[0].t1) slots = P.nxf + 1;
    return (size_t)kTileBytes + (size_t)slots * ((size_t)16 << P.log_r) + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D view of the source buffer: [word][row within set][set]; box = one tile row strip of <= 256 rows.
static bool make_tensor_map(const PassParams& P, CUtensorMap* map)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const uint32_t R = 1u << P.log_r, Wt = 16384u >> P.log_r;
    if (Wt > 256) return false;
    const cuuint64_t row_bytes = (cuuint64_t)P.pitch4 * 16;
    cuuint64_t gdim[3] = {(cuuint64_t)P.s4 * 4, R, P.nsets};
    cuuint64_t gstr[2] = {(cuuint64_t)P.src_row_stride * row_bytes, P.nsets > 1 ? (cuuint64_t)P.src_set_stride * row_bytes : (cuuint64_t)P.src_row_stride * row_bytes * R};
    cuuint32_t box[3] = {Wt, R < 256u ? R : 256u, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (gstr[0] >= (1ull << 40) || gstr[1] >= (1ull << 40)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)P.src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t dual_smem_bytes(const PassParams& P)
{
    return (size_t)3 * kTileBytes + (size_t)(P.xf[0].t1 ? 4 : 1) * ((size_t)16 << P.log_r) + 10 * sizeof(uint64_t);
}
template <int LR, int SHARD>
static cudaError_t launch_dual_inst(const PassParams& P, const CUtensorMap& map, unsigned grid, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_dual_kernel<LR, SHARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    ntt_pass_dual_kernel<LR, SHARD><<<grid, 2 * kThreads, dual_smem_bytes(P), stream>>>(P, map);
    return cudaGetLastError();
}

template <int LR, int NXF, int TMA>
static cudaError_t launch_inst(const PassParams& P, const CUtensorMap& map, unsigned grid, cudaStream_t stream)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<LR, NXF, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kTileBytes + 2 * NXF * (16 << LR) + 16);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ntt_pass_kernel<LR, NXF, TMA>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    ntt_pass_kernel<LR, NXF, TMA><<<grid, kThreads, pass_smem_bytes(P), stream>>>(P, map);
    return cudaGetLastError();
}

// name of the instantiation a pass is dispatched to, for the per-kernel lines of the benchmark (static storage)
static const char* kernel_name(bool dual, uint32_t lr, uint32_t a, uint32_t b)
{
    static char names[2][6][3][3][48];
    char* n = names[dual ? 1 : 0][lr - 5][a][b];
    if (!n[0]) { if (dual) snprintf(n, 48, "ntt_pass_dual_kernel<%u,%u>", lr, a); else snprintf(n, 48, "ntt_pass_kernel<%u,%u,%u>", lr, a, b); }
    return n;
}

cudaError_t launch_pass(const PassParams& Pin, int num_sms, cudaStream_t stream, const char** kernel)
{
    PassParams P = Pin;
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const unsigned long long nitems = (unsigned long long)P.nsets * groups;
    static const int ctas_per_sm = getenv("FASTECC_B200_CTAS_PER_SM") ? atoi(getenv("FASTECC_B200_CTAS_PER_SM")) : 2;   // tuning knob
    unsigned long long grid = (unsigned long long)num_sms * (ctas_per_sm > 0 ? ctas_per_sm : 2);
    if (grid > nitems) grid = nitems;
    if (grid == 0) return cudaSuccess;
    const unsigned g = (unsigned)grid;
    if (!P.tables) return cudaErrorInvalidValue;
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    static const bool no_tma = getenv("FASTECC_B200_NO_TMA") != nullptr;
    const bool tma = !no_tma && P.log_r >= 6 && make_tensor_map(P, &map);
    if (P.log_g && !tma) return cudaErrorNotSupported;          // sharded stores exist only in the TMA instantiations
    // single-transform passes with enough tiles to keep both groups of every SM busy: the dual schedule (3 tile buffers / SM)
    static const bool no_dual = getenv("FASTECC_B200_KERNEL") && !strcmp(getenv("FASTECC_B200_KERNEL"), "cta");
    if (tma && !no_dual && P.nxf == 1 && dual_smem_bytes(P) <= 232448 && nitems * P.strips_per_item >= 4ull * num_sms) {
        const unsigned gd = (unsigned)((unsigned long long)num_sms < nitems ? (unsigned long long)num_sms : nitems);
        if (kernel) *kernel = kernel_name(true, P.log_r, P.log_g ? 1 : 0, 0);
        switch (P.log_r) {
            case 6: return P.log_g ? launch_dual_inst<6, 1>(P, map, gd, stream) : launch_dual_inst<6, 0>(P, map, gd, stream);
            case 7: return P.log_g ? launch_dual_inst<7, 1>(P, map, gd, stream) : launch_dual_inst<7, 0>(P, map, gd, stream);
            case 8: return P.log_g ? launch_dual_inst<8, 1>(P, map, gd, stream) : launch_dual_inst<8, 0>(P, map, gd, stream);
            case 9: return P.log_g ? launch_dual_inst<9, 1>(P, map, gd, stream) : launch_dual_inst<9, 0>(P, map, gd, stream);
            case 10: return P.log_g ? launch_dual_inst<10, 1>(P, map, gd, stream) : launch_dual_inst<10, 0>(P, map, gd, stream);
            default: break;
        }
    }
    if (kernel) *kernel = kernel_name(false, P.log_r, P.nxf, tma ? (P.log_g ? 2 : 1) : 0);
#define FECC_CASE(L) case L: \
        if (tma && P.log_g) return P.nxf == 2 ? launch_inst<L, 2, 2>(P, map, g, stream) : launch_inst<L, 1, 2>(P, map, g, stream); \
        if (tma) return P.nxf == 2 ? launch_inst<L, 2, 1>(P, map, g, stream) : launch_inst<L, 1, 1>(P, map, g, stream); \
        else     return P.nxf == 2 ? launch_inst<L, 2, 0>(P, map, g, stream) : launch_inst<L, 1, 0>(P, map, g, stream);
    switch (P.log_r) {
        FECC_CASE(5) FECC_CASE(6) FECC_CASE(7) FECC_CASE(8) FECC_CASE(9) FECC_CASE(10)
        default: return cudaErrorInvalidValue;
    }
#undef FECC_CASE
}

} // namespace fecc
