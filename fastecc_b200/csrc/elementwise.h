// Element-wise GF(P) helpers (see elementwise.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace fecc {
cudaError_t launch_gf_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, cudaStream_t st);
cudaError_t launch_gf_inv(const uint32_t* a, uint32_t* out, size_t n, cudaStream_t st);
// rows of pitch4 16-byte chunks, the first s4 chunks of a row are multiplied by consts[row]; results canonical
cudaError_t launch_row_scale(uint32_t* d, size_t n_rows, uint32_t s4, uint32_t pitch4, const uint32_t* consts, int num_sms, cudaStream_t st);
// cross-GPU stream barrier over peer-mapped flag arrays (each: 8 arrival flags + 1 error word = 9 words, zero-initialised)
cudaError_t launch_shard_barrier(uint32_t* const* host_peers, uint32_t n_ranks, uint32_t rank, uint32_t epoch, cudaStream_t st);
}
