// Byte blocks <-> GF(P) words: the base-4096 -> base-4095 digit recoding of GF.md:72-104 (see byte_recode.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace fecc {
// n_blocks blocks of 4*W bytes (contiguous) -> rows of pitch_words words, W + 1 of them meaningful.  W % 4 == 0, W <= 1024.
cudaError_t launch_bytes_to_gfp(const void* d_bytes, uint32_t* d_words, size_t n_blocks, uint32_t W, size_t pitch_words, cudaStream_t stream);
cudaError_t launch_gfp_to_bytes(const uint32_t* d_words, void* d_bytes, size_t n_blocks, uint32_t W, size_t pitch_words, cudaStream_t stream);
}
