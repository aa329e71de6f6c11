// CPU emulation of ntt_pass_kernel: runs the *same* per-thread step functions (ntt_tile.cuh) and the same plans
// (plan.h) thread-by-thread on the host and checks the result against the oracle (oracle/gfp_oracle.c).
// Used by tests/test_emulation.py (-m "not gpu"): catches index / twiddle / plan errors without a GPU.
// Build: nvcc -O2 -o emulate_tile emulate_tile.cu -I../fastecc_b200/csrc -I../oracle ../oracle/liboracle.so
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "plan.h"
#include "emulate_pass.h"
#include "gfp_oracle.h"

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
