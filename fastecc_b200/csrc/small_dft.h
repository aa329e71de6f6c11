#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace fecc {
cudaError_t launch_small_dft(uint32_t* data, uint32_t pitch4, uint32_t s4, uint32_t N, uint32_t z, int encode,
                             uint32_t q, uint4 invN, const uint4* tw, cudaStream_t stream);
cudaError_t launch_pack  (const uint32_t* src, size_t src_pitch, uint32_t* dst, size_t dst_pitch, size_t n_rows, uint32_t size, cudaStream_t stream);
cudaError_t launch_unpack(const uint32_t* src, size_t src_pitch, uint32_t* dst, size_t dst_pitch, size_t n_rows, uint32_t size, cudaStream_t stream);
}
