// Helpers of the backward entry points shared between translation units (backward.cu, spectral_conv.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfno {

// gb[o] = sum_{b,p} gy[b][o][p]
int launch_bias_grad(const float* gy, int B, int C, int64_t hw, float* gb, cudaStream_t st);

}  // namespace sfno
