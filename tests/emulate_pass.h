// CPU emulation of ntt_pass_kernel (test infrastructure): runs the per-thread step functions of ntt_tile.cuh for
// tid = 0..255 in turn on host buffers.  Shared by emulate_tile.cu (plans vs oracle) and emulate_lib.cu (ctypes).
#pragma once
#include <vector>
#include <algorithm>
#include "plan.h"
#include "ntt_warp.cuh"
using namespace fecc;

static void emulate_pass(PassParams P)
{
    const uint32_t R = 1u << P.log_r;
    // per-set stage tables, as build_tables_kernel writes them
    const uint32_t nst = table_sets(P);
    std::vector<uint4> tables((size_t)nst * P.nxf * R);
    for (uint32_t set = 0; set < nst; ++set) for (uint32_t x = 0; x < P.nxf; ++x) for (uint32_t idx = 0; idx < R; ++idx)
        tables[((size_t)set * P.nxf + x) * R + idx] = idx ? P.tw[table_entry_exponent(P, x, set, idx)] : make_uint4(0, 0, 0, 0);
    P.tables = tables.data();
    P.table_set_stride = nst > 1 ? (P.nxf << P.log_r) : 0u;

    std::vector<uint4> tile(kTileChunks), tabs(2 * R);
    const uint32_t groups = (P.nstrips + P.strips_per_item - 1) / P.strips_per_item;
    const uint32_t nitems = P.nsets * groups;
    const uint32_t nsteps = num_steps(P.log_r, P.nxf);
    std::vector<RoundRegs> regs(kThreads);
    for (uint32_t item = 0; item < nitems; ++item) {
        const uint32_t set = item / groups, sg = item - set * groups;
        const uint32_t strip0 = sg * P.strips_per_item;
        const uint32_t strip1 = std::min(strip0 + P.strips_per_item, P.nstrips);
        for (uint32_t strip = strip0; strip < strip1; ++strip) {
            for (auto& c : tile) c = make_uint4(0xDEADBEEF, 0xDEADBEEF, 0xDEADBEEF, 0xDEADBEEF);
            for (uint32_t tid = 0; tid < kThreads; ++tid) load_tile_cpasync(P, set, strip, tid, tile.data());
            if (strip == strip0)
                for (uint32_t tid = 0; tid < kThreads; ++tid) load_tables_cpasync(P, set, tid, tabs.data());
            for (uint32_t s = 0; s < nsteps; ++s) {
                const Step st = step_of(P.log_r, P.nxf, s);
                const bool last = s + 1 == nsteps;
                for (uint32_t tid = 0; tid < kThreads; ++tid)
                    if (thread_active(P, tid, strip)) round_read(P, st.k, st.xfi == 0, tid, tile.data(), regs[tid]);
                for (uint32_t tid = 0; tid < kThreads; ++tid) {
                    if (!thread_active(P, tid, strip)) continue;
                    round_math(P, st, tid, set, tabs.data(), tabs.data() + R, regs[tid], 0);
                    if (last) round_write_global(P, st, tid, set, strip, regs[tid]);
                    else      round_write_tile(P, st.k, st.xfi == 0, tid, tile.data(), regs[tid]);
                }
            }
        }
    }
}


// Warp-private schedule (ntt_warp.cuh / ntt_pass_warp_kernel): TMA load into the swizzled tile, the nine phases of
// warp_phase() (each run for all 256 threads before the next: a superset of the kernel's __syncwarp ordering), TMA store.
static void emulate_pass_warp(PassParams P)
{
    const uint32_t LR = P.log_r, R = 1u << LR, Wt = 16384u >> LR;
    const uint32_t nst = table_sets(P);
    std::vector<uint4> tables((size_t)nst * P.nxf * R);
    for (uint32_t set = 0; set < nst; ++set) for (uint32_t x = 0; x < P.nxf; ++x) for (uint32_t idx = 0; idx < R; ++idx)
        tables[((size_t)set * P.nxf + x) * R + idx] = idx ? P.tw[table_entry_exponent(P, x, set, idx)] : make_uint4(0, 0, 0, 0);
    std::vector<uint4> tile(kTileChunks);
    uint2* t2 = reinterpret_cast<uint2*>(tile.data());
    std::vector<RoundRegs> regs(kThreads);
    std::vector<WarpAddr> wa(kThreads);
    for (uint32_t tid = 0; tid < kThreads; ++tid) wa[tid] = warp_addr(LR, warp_pos(LR, tid));
    const uint2* src2 = reinterpret_cast<const uint2*>(P.src);
    uint2* dst2 = reinterpret_cast<uint2*>(P.dst);
    const bool two = LR > (uint32_t)kStages;
    for (uint32_t set = 0; set < P.nsets; ++set) {
        const uint4* tw0 = tables.data() + (size_t)(nst > 1 ? set : 0) * P.nxf * R;
        const uint4* tw1 = tw0 + R;
        for (uint32_t strip = 0; strip < P.nstrips; ++strip) {
            for (uint32_t row = 0; row < R; ++row) for (uint32_t q2 = 0; q2 < Wt / 2; ++q2) {          // TMA load, OOB -> 0
                const uint32_t g2 = strip * (Wt / 2) + q2;
                const size_t grow = (size_t)set * P.src_set_stride + (size_t)row * P.src_row_stride;
                t2[cell8(LR, row, q2)] = (g2 >> 1) < P.s4 ? src2[grow * P.pitch4 * 2 + g2] : make_uint2(0, 0);
            }
            auto run = [&](int ph) {
                for (uint32_t tid = 0; tid < kThreads; ++tid)
                    if (warp_col_active(P, warp_pos(LR, tid), strip)) warp_phase(P, ph, tid, set, wa[tid], tile.data(), tw0, tw1, regs[tid], 0);
            };
            run(0); run(1);
            if (two) { run(2); run(3); run(4); if (P.nxf == 2) { run(5); run(6); run(7); } }
            run(8);
            for (uint32_t row = 0; row < R; ++row) for (uint32_t q2 = 0; q2 < Wt / 2; ++q2) {          // TMA store, clipped
                const uint32_t g2 = strip * (Wt / 2) + q2;
                if ((g2 >> 1) >= P.s4) continue;
                const size_t grow = (size_t)set * P.dst_set_stride + (size_t)row * P.dst_row_stride;
                dst2[grow * P.pitch4 * 2 + g2] = t2[cell8(LR, row, q2)];
            }
        }
    }
}
