// CUDA-core implicit-GEMM convolution contractions (fp32 accumulate): the fp32 arithmetic mode, and the bf16
// mode for layer shapes the tensor-core path does not take (3-channel image layers, odd channel counts).
#pragma once
#include <cuda_runtime.h>

#include "../../include/vaegan_b200.h"

namespace vg {
// Weights: dtype == VG_F32 -> fp32 master w[small_c][big_c][k][k];
//          dtype == VG_BF16 -> packed bf16 (down: wd[tap][small_c][big_c], up: wu[tap][big_c][small_c]).
int simt_conv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
                   int out_f32, cudaStream_t stream);
int simt_conv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, cudaStream_t stream);
int simt_conv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw,
                    cudaStream_t stream);
}  // namespace vg
