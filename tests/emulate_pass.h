// CPU emulation of ntt_pass_kernel (test infrastructure): runs the per-thread step functions of ntt_tile.cuh for
// tid = 0..255 in turn on host buffers.  Shared by emulate_tile.cu (plans vs oracle) and emulate_lib.cu (ctypes).
#pragma once
#include <vector>
#include <algorithm>
#include "plan.h"
using namespace fecc;

static void emulate_pass(PassParams P)
{
    const uint32_t R = 1u << P.log_r;
    // per-set stage tables, as build_tables_kernel writes them
    const uint32_t nst = table_sets(P);
    std::vector<uint4> tables((size_t)nst * P.nxf * R);
    for (uint32_t set = 0; set < nst; ++set) for (uint32_t x = 0; x < P.nxf; ++x) for (uint32_t idx = 0; idx < R; ++idx)
        tables[((size_t)set * P.nxf + x) * R + idx] = idx ? stage_entry(P.tw[table_entry_exponent(P, x, set, idx)]) : make_uint4(0, 0, 0, 0);
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
                    round_math(P, st, tid, set, tabs.data(), tabs.data() + R, regs[tid]);
                    if (last) round_write_global(P, st, tid, set, strip, regs[tid]);
                    else      round_write_tile(P, st.k, st.xfi == 0, tid, tile.data(), regs[tid]);
                }
            }
        }
    }
}

