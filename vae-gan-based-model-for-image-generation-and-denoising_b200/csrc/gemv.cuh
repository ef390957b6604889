// Full-extent single-output convolution (the discriminator's final Conv2d(8*ndf, 1, 4, 1, 0) on a 4x4 map,
// gan_code.py:84): per-sample dot products.  HBM-bound; warp-shuffle reductions, vectorised loads.
#pragma once
#include <cuda_runtime.h>

#include "../../include/vaegan_b200.h"

namespace vg {
// true when the geometry is "k == big_h == big_w, stride 1, pad 0, small 1x1x1"
bool is_gemv(const VgConvGeom* g);
// weights: dtype == VG_F32 -> fp32 master [1][bc][k][k]; VG_BF16 -> packed bf16 (down: wd[tap][1][bc]; up: wu[tap][bc][1])
int gemv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
              int out_f32, cudaStream_t st);
int gemv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, cudaStream_t st);
// dgrad with the BatchNorm-backward reduction of the layer below fused in (bf16, VG_EPI_BN_BWD)
bool gemv_up_fused_ok(const VgConvGeom* g, const VgEpilogue* ep);
int gemv_up_fused(const VgConvGeom* g, const void* small, const void* w, void* big, const VgEpilogue* ep, cudaStream_t st);
int gemv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, cudaStream_t st);
}  // namespace vg
