// ctypes-loadable CPU emulation of the pass kernel, used by the world_size-2 gloo test of the sharded encoder
// (tests/test_sharded_gloo.py): the same plan_encode_shard() descriptors the GPU path launches, executed on the host.
// Build: nvcc -O2 -shared -Xcompiler -fPIC -o libemulate.so emulate_lib.cu -I../fastecc_b200/csrc
#include "emulate_pass.h"

static std::vector<gf::Tw>& power_table()
{
    static std::vector<gf::Tw> t;
    if (t.empty()) { t.resize(kM); fill_power_table(t.data()); }
    return t;
}

extern "C" int emu_rs_encode_shard_pass(uint32_t* x, size_t N, int n_ranks, int rank, size_t size, size_t pitch, int which)
{
    if (!shard_supported(N, (uint32_t)n_ranks) || pitch % 4 || which < 0 || which > 2) return -1;
    Buffers b{x, nullptr, reinterpret_cast<const uint4*>(power_table().data()), (uint32_t)pitch, (uint32_t)size};
    emulate_pass(plan_encode_shard(b, N, (uint32_t)n_ranks, (uint32_t)rank, which));
    return 0;
}
extern "C" int emu_rs_encode_shard_pass_p2p(const uint32_t* src, uint32_t* const* peers, size_t N, int n_ranks, int rank, size_t size, size_t pitch, int which)
{
    if (!shard_p2p_supported(N, (uint32_t)n_ranks) || pitch % 4 || which < 0 || which > 2) return -1;
    emulate_pass(plan_encode_shard_p2p(src, peers, reinterpret_cast<const uint4*>(power_table().data()), (uint32_t)pitch, (uint32_t)size,
                                       N, (uint32_t)n_ranks, (uint32_t)rank, which));
    return 0;
}
extern "C" int emu_ntt_shard_pass_p2p(const uint32_t* src, uint32_t* const* peers, size_t N, int n_ranks, int rank, size_t size, size_t pitch, int inverse, int which)
{
    if (!ntt_shard_p2p_supported(N, (uint32_t)n_ranks) || pitch % 4 || which < 0 || which > 1) return -1;
    emulate_pass(plan_ntt_shard_p2p(src, peers, reinterpret_cast<const uint4*>(power_table().data()), (uint32_t)pitch, (uint32_t)size,
                                    N, (uint32_t)n_ranks, (uint32_t)rank, inverse != 0, which));
    return 0;
}
extern "C" int emu_rs_encode_asym(uint32_t* x, uint32_t* y, size_t N, size_t M, size_t size, size_t pitch)
{
    if (!asym_native(N, M)) return -1;
    Buffers b{x, y, reinterpret_cast<const uint4*>(power_table().data()), (uint32_t)pitch, (uint32_t)size};
    for (auto& p : plan_encode_asym(b, N, M)) emulate_pass(p);
    return 0;
}
extern "C" int emu_rs_encode(uint32_t* x, size_t N, size_t size, size_t pitch)
{
    Buffers b{x, nullptr, reinterpret_cast<const uint4*>(power_table().data()), (uint32_t)pitch, (uint32_t)size};
    for (auto& p : plan_encode(b, N)) emulate_pass(p);
    return 0;
}
