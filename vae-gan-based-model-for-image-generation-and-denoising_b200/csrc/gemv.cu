#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "pdl.cuh"
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
    pdl_enter();
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
    pdl_enter();
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
    pdl_enter();
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

// ---- bf16 fast paths (packed weights [tap][c] have the activation's own linear order; no padded channels):
// 16-byte vectors of 8 bf16 everywhere.
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}

// one block per sample: logit[b] = x[b] . w  over K = kk*c elements
__global__ void __launch_bounds__(256) gemv_down_bf16_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ w,
                                                            const float* __restrict__ bias, void* __restrict__ out,
                                                            int out_f32, int kvec) {
    pdl_enter();
    __shared__ float red[8];
    const uint4* xb = reinterpret_cast<const uint4*>(x) + static_cast<long long>(blockIdx.x) * kvec;
    const uint4* wv = reinterpret_cast<const uint4*>(w);
    float acc = 0.f;
    for (int i = threadIdx.x; i < kvec; i += blockDim.x) {
        float a[8], b[8];
        unpack8(__ldg(xb + i), a);
        unpack8(__ldg(wv + i), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(a[j], b[j], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        if (bias != nullptr) t += bias[0];
        if (out_f32) static_cast<float*>(out)[blockIdx.x] = t;
        else static_cast<__nv_bfloat16*>(out)[blockIdx.x] = __float2bfloat16_rn(t);
    }
}

// dx[b][i] = s[b] * w[i]
__global__ void __launch_bounds__(256) gemv_up_bf16_kernel(const __nv_bfloat16* __restrict__ s,
                                                          const __nv_bfloat16* __restrict__ w,
                                                          __nv_bfloat16* __restrict__ dx, int batch, int kvec) {
    pdl_enter();
    const long long total = static_cast<long long>(batch) * kvec;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / kvec), k = static_cast<int>(i - static_cast<long long>(b) * kvec);
        const float sb = __bfloat162float(s[b]);
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(w) + k), a);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(sb * a[2 * j], sb * a[2 * j + 1]);
            o[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(dx)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// dw[c][tap] += sum_b s[b] * x[b][tap][c]: thread = (8-element column vector, batch chunk); chunks meet in fp32 atomics
__global__ void __launch_bounds__(256) gemv_wgrad_bf16_kernel(const __nv_bfloat16* __restrict__ s,
                                                             const __nv_bfloat16* __restrict__ x,
                                                             float* __restrict__ dw, int batch, int kk, int c, int kvec,
                                                             int chunk) {
    pdl_enter();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kvec) return;
    const int b0 = blockIdx.y * chunk, b1 = min(batch, b0 + chunk);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const uint4* xv = reinterpret_cast<const uint4*>(x) + k;
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
        float a[8];
        unpack8(__ldg(xv + static_cast<long long>(b) * kvec), a);
        const float sb = __bfloat162float(s[b]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(sb, a[j], acc[j]);
    }
    const int i0 = k * 8, tap = i0 / c, c0 = i0 - tap * c;     // c is a multiple of 8: a vector stays inside one tap
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dw + static_cast<long long>(c0 + j) * kk + tap, acc[j]);
}

// dgrad of the single-output head with the BatchNorm backward of the layer below folded in (VG_EPI_BN_BWD):
//   dz[b][tap][c] = s[b] * w[tap][c] * act'(x*scale + shift),  sums[g][0][c] += dz,  sums[g][1][c] += dz * xhat
// thread = (8-channel vector of one tap, chunk of the batch); a block holds 256/(c/8) taps of every channel vector,
// combines them in shared memory and issues one atomic per (channel, quantity).
constexpr int kGemvChunk = 16;
__global__ void __launch_bounds__(256) gemv_up_bnbwd_kernel(const __nv_bfloat16* __restrict__ s,
                                                           const __nv_bfloat16* __restrict__ w,
                                                           const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ stats, float* __restrict__ sums,
                                                           __nv_bfloat16* __restrict__ dz, int batch, int c, int kvec,
                                                           int group_batch, float neg_slope) {
    pdl_enter();
    __shared__ float red[256][17];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;           // kvec is a multiple of 256 (checked by the launcher)
    const int b0 = blockIdx.y * kGemvChunk, b1 = min(batch, b0 + kGemvChunk);
    const int grp = b0 / group_batch;
    const int cv = c / 8, c0 = (k % cv) * 8;
    const float* st = stats + static_cast<long long>(grp) * 4 * c;
    float mean[8], rstd[8], sc[8], sh[8], wv[8], a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mean[j] = __ldg(st + c0 + j);
        rstd[j] = __ldg(st + c + c0 + j);
        sc[j] = __ldg(st + 2 * c + c0 + j);
        sh[j] = __ldg(st + 3 * c + c0 + j);
        a0[j] = a1[j] = 0.f;
    }
    unpack8(__ldg(reinterpret_cast<const uint4*>(w) + k), wv);
    // the batch loop used to issue ONE 16-byte load per iteration and wait for it (16 dependent round trips per
    // thread on a 128-block grid: 17 us for 17 MB); eight saved-tensor vectors are now in flight per thread
    constexpr int PF = 8;
    for (int bb = b0; bb < b1; bb += PF) {
      uint4 xr[PF];
#pragma unroll
      for (int u = 0; u < PF; ++u)
          if (bb + u < b1) xr[u] = __ldg(reinterpret_cast<const uint4*>(x) + static_cast<long long>(bb + u) * kvec + k);
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int b = bb + u;
        if (b >= b1) break;
        const long long i = static_cast<long long>(b) * kvec + k;
        float xv[8];
        unpack8(xr[u], xv);
        const float sb = __bfloat162float(s[b]);
        uint32_t o[4];
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z = fmaf(xv[j], sc[j], sh[j]);
            d[j] = sb * wv[j] * (z > 0.f ? 1.f : neg_slope);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(d[2 * j], d[2 * j + 1]);
            o[j] = *reinterpret_cast<uint32_t*>(&h);
            d[2 * j] = __uint_as_float(o[j] << 16);                 // statistics of exactly what is stored
            d[2 * j + 1] = __uint_as_float(o[j] & 0xFFFF0000u);
        }
        reinterpret_cast<uint4*>(dz)[i] = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0[j] += d[j];
            a1[j] = fmaf(d[j], (xv[j] - mean[j]) * rstd[j], a1[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = a0[j]; red[threadIdx.x][8 + j] = a1[j]; }
    __syncthreads();
    if (threadIdx.x < cv) {          // the block's threads t, t+cv, t+2cv, ... hold the same channels for different taps
        float* g0 = sums + static_cast<long long>(grp) * 2 * c;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float t = 0.f;
            for (int r = threadIdx.x; r < 256; r += cv) t += red[r][j];
            atomicAdd(g0 + (j < 8 ? 0 : c) + c0 + (j & 7), t);
        }
    }
}

int valid_c(const VgConvGeom* g) { return g->big_c_valid > 0 ? g->big_c_valid : g->big_c; }

bool bf16_fast(const VgConvGeom* g, const void* a, const void* b) {
    return valid_c(g) == g->big_c && g->big_c % 8 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

}  // namespace

bool is_gemv(const VgConvGeom* g) {
    return g->small_c == 1 && g->small_h == 1 && g->small_w == 1 && g->stride == 1 && g->pad == 0 &&
           g->kernel == g->big_h && g->kernel == g->big_w;
}

int gemv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
              int out_f32, cudaStream_t stm) {
    const int kk = g->kernel * g->kernel;
    if (dtype == VG_BF16 && bf16_fast(g, big, w))
        launch_k(gemv_down_bf16_kernel, dim3(g->batch), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(big),
                                                         static_cast<const __nv_bfloat16*>(w), bias, small, out_f32,
                                                         kk * g->big_c / 8);
    else if (dtype == VG_BF16)
        launch_k(gemv_down_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(g->batch), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(big), static_cast<const __nv_bfloat16*>(w), bias, small, out_f32, kk,
            g->big_c, valid_c(g));
    else
        launch_k(gemv_down_kernel<float, float>, dim3(g->batch), dim3(256), 0, stm, static_cast<const float*>(big),
                                                                  static_cast<const float*>(w), bias, small, out_f32,
                                                                  kk, g->big_c, valid_c(g));
    VG_LAUNCHED();
    return VG_OK;
}

int gemv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, cudaStream_t stm) {
    const int kk = g->kernel * g->kernel;
    const long long total = static_cast<long long>(g->batch) * kk * g->big_c;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
    if (dtype == VG_BF16 && bf16_fast(g, big, w)) {
        const int vblocks = static_cast<int>(std::min<long long>((total / 8 + 255) / 256, 148 * 8));
        launch_k(gemv_up_bf16_kernel, dim3(vblocks), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(small),
                                                      static_cast<const __nv_bfloat16*>(w),
                                                      static_cast<__nv_bfloat16*>(big), g->batch, kk * g->big_c / 8);
    } else if (dtype == VG_BF16)
        launch_k(gemv_up_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(blocks), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(small), static_cast<const __nv_bfloat16*>(w),
            static_cast<__nv_bfloat16*>(big), g->batch, kk, g->big_c, valid_c(g));
    else
        launch_k(gemv_up_kernel<float, float>, dim3(blocks), dim3(256), 0, stm, static_cast<const float*>(small),
                                                              static_cast<const float*>(w), static_cast<float*>(big),
                                                              g->batch, kk, g->big_c, valid_c(g));
    VG_LAUNCHED();
    return VG_OK;
}

bool gemv_up_fused_ok(const VgConvGeom* g, const VgEpilogue* ep) {
    if (ep == nullptr || ep->mode != VG_EPI_BN_BWD || !is_gemv(g)) return false;
    const int c = g->big_c, kvec = g->kernel * g->kernel * c / 8, groups = ep->groups > 0 ? ep->groups : 1;
    if (valid_c(g) != c || c % 8 != 0 || c / 8 > 256 || 256 % (c / 8) != 0 || kvec % 256 != 0 || ep->channels != c) return false;
    if (g->batch % groups != 0 || (g->batch / groups) % kGemvChunk != 0) return false;
    if (ep->act != VG_ACT_NONE && ep->act != VG_ACT_RELU && ep->act != VG_ACT_LEAKY) return false;
    return ep->sums != nullptr && ep->stats != nullptr && ep->x != nullptr;
}

int gemv_up_fused(const VgConvGeom* g, const void* small, const void* w, void* big, const VgEpilogue* ep,
                  cudaStream_t stm) {
    const int c = g->big_c, kvec = g->kernel * g->kernel * c / 8, groups = ep->groups > 0 ? ep->groups : 1;
    const float neg_slope = ep->act == VG_ACT_RELU ? 0.f : (ep->act == VG_ACT_LEAKY ? ep->slope : 1.f);
    const dim3 grid(kvec / 256, (g->batch + kGemvChunk - 1) / kGemvChunk);
    launch_k(gemv_up_bnbwd_kernel, dim3(grid), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(small),
                                                static_cast<const __nv_bfloat16*>(w),
                                                static_cast<const __nv_bfloat16*>(ep->x), ep->stats, ep->sums,
                                                static_cast<__nv_bfloat16*>(big), g->batch, c, kvec, g->batch / groups,
                                                neg_slope);
    VG_LAUNCHED();
    return VG_OK;
}

int gemv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, cudaStream_t stm) {
    const int kk = g->kernel * g->kernel;
    const long long K = static_cast<long long>(kk) * g->big_c;
    const int blocks = static_cast<int>((K + 255) / 256);
    if (dtype == VG_BF16 && bf16_fast(g, big, big)) {
        const int kvec = static_cast<int>(K / 8), chunk = 16;
        const dim3 grid((kvec + 255) / 256, (g->batch + chunk - 1) / chunk);
        launch_k(gemv_wgrad_bf16_kernel, dim3(grid), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(small),
                                                      static_cast<const __nv_bfloat16*>(big), dw, g->batch, kk, g->big_c,
                                                      kvec, chunk);
    } else if (dtype == VG_BF16)
        launch_k(gemv_wgrad_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stm, static_cast<const __nv_bfloat16*>(small),
                                                                  static_cast<const __nv_bfloat16*>(big), dw, g->batch,
                                                                  kk, g->big_c, valid_c(g));
    else
        launch_k(gemv_wgrad_kernel<float>, dim3(blocks), dim3(256), 0, stm, static_cast<const float*>(small),
                                                          static_cast<const float*>(big), dw, g->batch, kk, g->big_c,
                                                          valid_c(g));
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace vg
