// Order-3 / order-9 step of the N = 3 * 2^k and 9 * 2^k transforms (see mixed_radix.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace fecc {
// rows r*k2 + n1 of src (k2 < M, n1 < r)  ->  rows k2 + M*k1 of dst; s4 16-byte chunks per row, rows pitch4 chunks apart
cudaError_t launch_radix_pass(const uint32_t* src, uint32_t* dst, uint32_t pitch4, uint32_t s4, uint32_t r, uint32_t M, bool inverse, const uint4* tw, int num_sms, cudaStream_t st);
}
