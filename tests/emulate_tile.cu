// CPU emulation of ntt_pass_kernel: runs the *same* per-thread step functions (ntt_tile.cuh) and the same plans
// (plan.h) thread-by-thread on the host and checks the result against the oracle (oracle/gfp_oracle.c).
// Used by tests/test_emulation.py (-m "not gpu"): catches index / twiddle / plan errors without a GPU.
// Build: nvcc -O2 -o emulate_tile emulate_tile.cu -I../fastecc_b200/csrc -I../oracle ../oracle/liboracle.so
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "plan.h"
#include "gfp_oracle.h"

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

static std::vector<gf::Tw> g_tw;

static int check(const char* what, size_t N, size_t S, int mode /*0 fwd,1 inv,2 encode*/)
{
    const size_t pitch = (S + 3) / 4 * 4;
    std::vector<uint32_t> x(N * pitch, 0), y(N * pitch, 0), ref(N * S);
    oracle_fill_B(ref.data(), N * S);
    for (size_t i = 0; i < N; i++) memcpy(&x[i * pitch], &ref[i * S], S * 4);
    Buffers b{x.data(), y.data(), reinterpret_cast<const uint4*>(g_tw.data()), (uint32_t)pitch, (uint32_t)S};
    std::vector<PassParams> plan = (mode == 2) ? plan_encode(b, N) : plan_ntt(b, N, mode == 1);
    for (auto& p : plan) emulate_pass(p);
    if (mode == 2) oracle_rs_encode(ref.data(), N, S); else oracle_ntt(ref.data(), N, S, mode);
    size_t bad = 0;
    for (size_t i = 0; i < N; i++) for (size_t k = 0; k < S; k++) if (x[i * pitch + k] != ref[i * S + k]) { if (!bad) printf("   first mismatch row %zu word %zu: got %u want %u\n", i, k, x[i*pitch+k], ref[i*S+k]); bad++; }
    printf("%-8s N=2^%-2u S=%-5zu passes=%zu : %s (%zu mismatches)\n", what, ilog2(N), S, plan.size(), bad ? "FAIL" : "ok", bad);
    return bad != 0;
}

int main(int argc, char** argv)
{
    g_tw.resize(kM);
    fill_power_table(g_tw.data());
    int fails = 0;
    const bool big = argc > 1 && !strcmp(argv[1], "big");
    struct Cfg { unsigned ln; size_t s; };
    std::vector<Cfg> cfgs = { {5, 8}, {5, 1024}, {6, 1024}, {7, 1024}, {8, 20}, {9, 36}, {10, 16}, {10, 24}, {11, 16}, {12, 8}, {13, 4}, {7, 513} };
    if (big) { cfgs.push_back({16, 16}); cfgs.push_back({19, 16}); cfgs.push_back({20, 4}); cfgs.push_back({15, 32}); }
    for (auto c : cfgs) {
        const size_t N = (size_t)1 << c.ln;
        fails += check("ntt-fwd", N, c.s, 0);
        fails += check("ntt-inv", N, c.s, 1);
        if (c.ln <= 19) fails += check("encode", N, c.s, 2);
    }
    printf("%s\n", fails ? "EMULATION FAILED" : "EMULATION OK");
    return fails ? 1 : 0;
}
