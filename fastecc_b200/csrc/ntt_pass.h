#pragma once
#include <cuda_runtime.h>
#include "ntt_tile.cuh"
namespace fecc {
size_t      pass_smem_bytes(const PassParams& P);
cudaError_t launch_pass(const PassParams& P, int num_sms, cudaStream_t stream);
}
