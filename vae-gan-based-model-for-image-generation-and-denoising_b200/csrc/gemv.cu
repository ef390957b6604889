#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "gemv.cuh"

namespace vg {
namespace {

__device__ __forceinline__ float ld(const float* p, long long i) { return __ldg(p + i); }
__device__ __forceinline__ float ld(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void st(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void st(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

// weight element for (tap, channel): fp32 master is [c][tap], packed bf16 is [tap][c_pad]
template <typename WT>
__device__ __forceinline__ float wat(const WT* w, int tap, int c, int kk, int cpad) {
    if constexpr (sizeof(WT) == 4) return ld(w, static_cast<long long>(c) * kk + tap);
    else return ld(w, static_cast<long long>(tap) * cpad + c);
}

// one block per sample: logit[b] = sum_{tap,c} x[b][tap][c] * w(tap, c)
template <typename T, typename WT>
__global__ void __launch_bounds__(256) gemv_down_kernel(const T* __restrict__ x, const WT* __restrict__ w,
                                                       const float* __restrict__ bias, void* __restrict__ out,
                                                       int out_f32, int kk, int cpad, int cvalid) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const long long K = static_cast<long long>(kk) * cpad;
    const T* xb = x + b * K;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < K; i += blockDim.x) {
        const int tap = static_cast<int>(i / cpad), c = static_cast<int>(i - static_cast<long long>(tap) * cpad);
        if (c < cvalid) acc = fmaf(ld(xb, i), wat(w, tap, c, kk, cpad), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        if (bias != nullptr) t += bias[0];
        if (out_f32) static_cast<float*>(out)[b] = t;
        else st(static_cast<T*>(out), b, t);
    }
}

// dx[b][tap][c] = s[b] * w(tap, c)
template <typename T, typename WT>
__global__ void __launch_bounds__(256) gemv_up_kernel(const T* __restrict__ s, const WT* __restrict__ w,
                                                     T* __restrict__ dx, int batch, int kk, int cpad, int cvalid) {
    const long long K = static_cast<long long>(kk) * cpad, total = K * batch;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / K);
        const long long r = i - b * K;
        const int tap = static_cast<int>(r / cpad), c = static_cast<int>(r - static_cast<long long>(tap) * cpad);
        st(dx, i, c < cvalid ? ld(s, b) * wat(w, tap, c, kk, cpad) : 0.f);
    }
}

// dw[c][tap] += sum_b s[b] * x[b][tap][c]   (master layout [1][c][k][k])
template <typename T>
__global__ void __launch_bounds__(256) gemv_wgrad_kernel(const T* __restrict__ s, const T* __restrict__ x,
                                                        float* __restrict__ dw, int batch, int kk, int cpad,
                                                        int cvalid) {
    const long long K = static_cast<long long>(kk) * cpad;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < K;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i / cpad), c = static_cast<int>(i - static_cast<long long>(tap) * cpad);
        if (c >= cvalid) continue;
        float acc = 0.f;
        for (int b = 0; b < batch; ++b) acc = fmaf(ld(s, b), ld(x, b * K + i), acc);
        dw[static_cast<long long>(c) * kk + tap] += acc;
    }
}

int valid_c(const VgConvGeom* g) { return g->big_c_valid > 0 ? g->big_c_valid : g->big_c; }

}  // namespace

bool is_gemv(const VgConvGeom* g) {
    return g->small_c == 1 && g->small_h == 1 && g->small_w == 1 && g->stride == 1 && g->pad == 0 &&
           g->kernel == g->big_h && g->kernel == g->big_w;
}

int gemv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
              int out_f32, cudaStream_t stm) {
    const int kk = g->kernel * g->kernel;
    if (dtype == VG_BF16)
        gemv_down_kernel<__nv_bfloat16, __nv_bfloat16><<<g->batch, 256, 0, stm>>>(
            static_cast<const __nv_bfloat16*>(big), static_cast<const __nv_bfloat16*>(w), bias, small, out_f32, kk,
            g->big_c, valid_c(g));
    else
        gemv_down_kernel<float, float><<<g->batch, 256, 0, stm>>>(static_cast<const float*>(big),
                                                                  static_cast<const float*>(w), bias, small, out_f32,
                                                                  kk, g->big_c, valid_c(g));
    VG_LAUNCHED();
    return VG_OK;
}

int gemv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, cudaStream_t stm) {
    const int kk = g->kernel * g->kernel;
    const long long total = static_cast<long long>(g->batch) * kk * g->big_c;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
    if (dtype == VG_BF16)
        gemv_up_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, stm>>>(
            static_cast<const __nv_bfloat16*>(small), static_cast<const __nv_bfloat16*>(w),
            static_cast<__nv_bfloat16*>(big), g->batch, kk, g->big_c, valid_c(g));
    else
        gemv_up_kernel<float, float><<<blocks, 256, 0, stm>>>(static_cast<const float*>(small),
                                                              static_cast<const float*>(w), static_cast<float*>(big),
                                                              g->batch, kk, g->big_c, valid_c(g));
    VG_LAUNCHED();
    return VG_OK;
}

int gemv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, cudaStream_t stm) {
    const int kk = g->kernel * g->kernel;
    const long long K = static_cast<long long>(kk) * g->big_c;
    const int blocks = static_cast<int>((K + 255) / 256);
    if (dtype == VG_BF16)
        gemv_wgrad_kernel<__nv_bfloat16><<<blocks, 256, 0, stm>>>(static_cast<const __nv_bfloat16*>(small),
                                                                  static_cast<const __nv_bfloat16*>(big), dw, g->batch,
                                                                  kk, g->big_c, valid_c(g));
    else
        gemv_wgrad_kernel<float><<<blocks, 256, 0, stm>>>(static_cast<const float*>(small),
                                                          static_cast<const float*>(big), dw, g->batch, kk, g->big_c,
                                                          valid_c(g));
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace vg
