// One Reed-Solomon encode of N blocks sharded over G GPUs, driven from C++ through the C ABI only (include/fastecc_b200.h):
// one process per GPU, the ranks exchange nothing but 3 x 64-byte CUDA IPC handles (here through files in a directory; any
// transport works).  Launch G copies:   shard_example <rank> <G> <log2 N> <SIZE_words> <exchange dir>
// Global block l*G + rank is local row l, for the data going in and the parity coming out.  The reference has no multi-GPU path;
// what is sharded is the transpose between the four-step passes (TransposeMatrix, ntt.cpp:322-341).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <cuda_runtime.h>
#include "fastecc_b200.h"

#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "%s failed: %s\n", #x, fastecc_b200_last_error()); return 1; } } while (0)

static void put(const std::string& dir, int rank, int k, const void* h) { std::string p = dir + "/h" + std::to_string(rank) + "_" + std::to_string(k); FILE* f = fopen((p + ".tmp").c_str(), "wb"); fwrite(h, 1, 64, f); fclose(f); rename((p + ".tmp").c_str(), p.c_str()); }
static void get(const std::string& dir, int rank, int k, void* h) { std::string p = dir + "/h" + std::to_string(rank) + "_" + std::to_string(k); FILE* f; while (!(f = fopen(p.c_str(), "rb"))) usleep(1000); fread(h, 1, 64, f); fclose(f); }

int main(int argc, char** argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s rank G log2N SIZE_words dir\n", argv[0]); return 2; }
    const int rank = atoi(argv[1]), G = atoi(argv[2]);
    const size_t N = (size_t)1 << atoi(argv[3]), S = (size_t)atoi(argv[4]);
    const std::string dir = argv[5];
    CHECK(fastecc_b200_init(rank));
    const size_t rows = N / G, bytes = rows * S * 4;
    void* own[3] = {fastecc_b200_dev_alloc(bytes), fastecc_b200_dev_alloc(bytes), fastecc_b200_dev_alloc(4 * FASTECC_B200_BARRIER_WORDS)};
    cudaMemset(own[2], 0, 4 * FASTECC_B200_BARRIER_WORDS);
    cudaDeviceSynchronize();
    for (int k = 0; k < 3; ++k) { char h[64]; CHECK(fastecc_b200_ipc_export(own[k], h)); put(dir, rank, k, h); }
    std::vector<uint32_t*> peers[3];
    for (int k = 0; k < 3; ++k) for (int r = 0; r < G; ++r) {
        void* p = own[k];
        if (r != rank) { char h[64]; get(dir, r, k, h); CHECK(fastecc_b200_ipc_open(h, &p)); }
        peers[k].push_back((uint32_t*)p);
    }
    // data0[i] = i % P over the GLOBAL array (RS.cpp:28-29), this rank's rows
    std::vector<uint32_t> host(rows * S);
    for (size_t l = 0; l < rows; ++l) for (size_t k = 0; k < S; ++k) host[l * S + k] = (uint32_t)((((l * G + rank) * S) + k) % FASTECC_B200_P);
    cudaMemcpy(own[0], host.data(), bytes, cudaMemcpyHostToDevice);
    uint32_t epoch = 0;
    CHECK(fastecc_b200_shard_barrier(peers[2].data(), G, rank, ++epoch, nullptr));           // everybody has mapped everything
    CHECK(fastecc_b200_rs_encode_shard_p2p(peers[0].data(), peers[1].data(), peers[2].data(), &epoch, N, G, rank, S, S, nullptr));
    CHECK(fastecc_b200_shard_barrier(peers[2].data(), G, rank, ++epoch, nullptr));           // nobody is still storing into our X
    cudaMemcpy(host.data(), own[0], bytes, cudaMemcpyDeviceToHost);
    uint32_t h = 314159253u;                                                                   // main.cpp:203-212 over this rank's parity rows
    for (size_t i = 0; i < rows * S; ++i) h = (h + host[i]) * 123456791u + (h >> 17);
    printf("rank %d: parity rows %zu, local hash %u\n", rank, rows, h);
    return 0;
}
