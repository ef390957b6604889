// VAE-GAN loss-side kernels: reparameterisation + KL prior, BCE adversarial terms, MSE (pixel or feature)
// reconstruction, fused Adam, Philox normal noise.  fp32 arithmetic, vectorised coalesced access, warp-shuffle +
// shared-memory block reductions, deterministic two-stage grid reductions.
//   reference call sites: vaegan_code.py:75-78 (reparam), :99-101,:115 (BCE), :113 (MSE), :114 (KL), :117 (total),
//   :105,:134-135 (Adam), :77,:91-92 (randn_like).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "pdl.cuh"
#include "vec.cuh"

namespace vg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum over the block; result valid in thread 0.
template <typename F>
__device__ __forceinline__ F block_sum(F v, F* smem /* [32] */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    F r = 0;
    if (warp == 0) {
        r = lane < (blockDim.x + 31) / 32 ? smem[lane] : F(0);
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ---- reparameterisation + KL: one element per thread; block partials meet in a double accumulator and the block
// that arrives last (ticket) writes the KL term and re-arms both (one step at a time per process uses this kernel)
__device__ double g_kl_acc;
__device__ unsigned int g_kl_ticket;

template <typename T>
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                         const float* __restrict__ eps, int n, int batch,
                                                         T* __restrict__ z, float* __restrict__ kl_out) {
    pdl_enter();
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float m = mu[i];
        const float lv = fminf(10.f, fmaxf(-10.f, logvar[i]));
        const float sd = expf(0.5f * lv);
        const float zz = fmaf(sd, eps[i], m);
        if constexpr (sizeof(T) == 4) z[i] = zz; else z[i] = __float2bfloat16_rn(zz);
        acc += static_cast<double>(1.f + lv - m * m - expf(lv));
    }
    const double part = block_sum(acc, red);
    if (threadIdx.x == 0) {
        atomicAdd(&g_kl_acc, part);
        __threadfence();
        if (atomicAdd(&g_kl_ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            const double tot = atomicAdd(&g_kl_acc, 0.0);
            if (kl_out != nullptr) *kl_out = static_cast<float>(-0.5 * tot / batch);
            g_kl_acc = 0.0;
            g_kl_ticket = 0;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) reparam_bwd_kernel(const T* __restrict__ dz, const float* __restrict__ mu,
                                                              const float* __restrict__ logvar,
                                                              const float* __restrict__ eps, int n, int batch,
                                                              const float* __restrict__ kl_weight_ptr, float kl_weight,
                                                              float* __restrict__ dmu, float* __restrict__ dlogvar) {
    pdl_enter();
    const float w = (kl_weight_ptr ? *kl_weight_ptr : kl_weight) / batch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g;
        if constexpr (sizeof(T) == 4) g = dz[i]; else g = __bfloat162float(dz[i]);
        const float m = mu[i], lraw = logvar[i];
        const float lv = fminf(10.f, fmaxf(-10.f, lraw));
        const float pass = (lraw >= -10.f && lraw <= 10.f) ? 1.f : 0.f;
        const float sd = expf(0.5f * lv);
        dmu[i] = g + w * m;
        dlogvar[i] = pass * (g * 0.5f * sd * eps[i] + w * 0.5f * (expf(lv) - 1.f));
    }
}

// ---- BCE on probabilities (nn.BCELoss, mean reduction, log clamped at -100), loss and d(loss)/dp
__global__ void __launch_bounds__(1024) bce_kernel(const float* __restrict__ p, int n, float target, float weight,
                                                  float* __restrict__ loss_out, int accumulate,
                                                  float* __restrict__ dp) {
    pdl_enter();
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float pi = p[i];
        const float l1 = fmaxf(logf(pi), -100.f), l0 = fmaxf(log1pf(-pi), -100.f);
        acc += static_cast<double>(-(target * l1 + (1.f - target) * l0));
        if (dp != nullptr) dp[i] = weight / n * (pi - target) / fmaxf((1.f - pi) * pi, 1e-12f);
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0 && loss_out != nullptr) {
        const float l = static_cast<float>(tot / n);
        *loss_out = accumulate ? *loss_out + l : l;
    }
}

// ---- MSE(a, b) (mean) + gradient wrt a, optionally added to an incoming gradient.  Stage 1: per-block partials.
__global__ void __launch_bounds__(kThreads) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              long long n, float weight,
                                                              const float* __restrict__ grad_in,
                                                              float* __restrict__ grad_out,
                                                              double* __restrict__ partial) {
    pdl_enter();
    __shared__ double red[32];
    double acc = 0.0;
    const float gs = 2.f * weight / static_cast<float>(n);
    const long long nvec = n / 4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 av = reinterpret_cast<const float4*>(a)[i], bv = reinterpret_cast<const float4*>(b)[i];
        const float d0 = av.x - bv.x, d1 = av.y - bv.y, d2 = av.z - bv.z, d3 = av.w - bv.w;
        acc += static_cast<double>(d0 * d0 + d1 * d1) + static_cast<double>(d2 * d2 + d3 * d3);
        if (grad_out != nullptr) {
            float4 g = make_float4(gs * d0, gs * d1, gs * d2, gs * d3);
            if (grad_in != nullptr) {
                const float4 gi = reinterpret_cast<const float4*>(grad_in)[i];
                g.x += gi.x; g.y += gi.y; g.z += gi.z; g.w += gi.w;
            }
            reinterpret_cast<float4*>(grad_out)[i] = g;
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * 4 + threadIdx.x; i < n; i += blockDim.x) {
            const float d = a[i] - b[i];
            acc += static_cast<double>(d * d);
            if (grad_out != nullptr) grad_out[i] = gs * d + (grad_in ? grad_in[i] : 0.f);
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void mse_finalize_kernel(const double* partial, int blocks, long long n, float* loss_out) {
    pdl_enter();
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < blocks; i += blockDim.x) acc += partial[i];
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) *loss_out = static_cast<float>(tot / static_cast<double>(n));
}

// ---- one launch per discriminator update: BCE(D(real), real_label) + BCE(D(fake), fake_label) of the stacked
// probability vector p[2n] (real half first), both gradient seeds, the summed loss (vaegan_code.py:99-101)
__global__ void __launch_bounds__(1024) bce_pair_kernel(const float* __restrict__ p, int n, float t_real, float t_fake,
                                                       float weight, float* __restrict__ loss_out,
                                                       float* __restrict__ dp) {
    pdl_enter();
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
        const float target = i < n ? t_real : t_fake;
        const float pi = p[i];
        const float l1 = fmaxf(logf(pi), -100.f), l0 = fmaxf(log1pf(-pi), -100.f);
        acc += static_cast<double>(-(target * l1 + (1.f - target) * l0));
        if (dp != nullptr) dp[i] = weight / n * (pi - target) / fmaxf((1.f - pi) * pi, 1e-12f);
    }
    const double tot = block_sum(acc, red);      // (each half is a mean over n: the sum of both means = tot / n)
    if (threadIdx.x == 0 && loss_out != nullptr) *loss_out = static_cast<float>(tot / n);
}

// ---- MSE(a, b) (mean) with its gradient and, from the last block to finish, the step's total loss - one launch for
// vaegan_code.py:113 + :117 (no finalize / total kernels).  T = float (pixel MSE of the fp32 NCHW reconstruction) or
// bf16 (Dis_l feature matching on NHWC discriminator features, README.md:11-14); the gradient is written in T.
struct MseAcc {
    double acc;
    unsigned int ticket;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) mse_total_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                            long long n, float weight, const T* __restrict__ grad_in,
                                                            T* __restrict__ grad_out, float* __restrict__ loss_out,
                                                            const float* __restrict__ kl, const float* __restrict__ adv,
                                                            const float* __restrict__ w_kl_ptr, float w_adv,
                                                            float* __restrict__ total_out, MseAcc* __restrict__ ws) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    __shared__ double red[32];
    double acc = 0.0;
    const float gs = 2.f * weight / static_cast<float>(n);
    const long long nvec = n / V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float av[V], bv[V], gv[V];
        Vec<T>::load(a + i * V, av);
        Vec<T>::load(b + i * V, bv);
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float d = av[j] - bv[j];
            part = fmaf(d, d, part);
            gv[j] = gs * d;
        }
        acc += static_cast<double>(part);
        if (grad_out != nullptr) {
            if (grad_in != nullptr) {
                float gi[V];
                Vec<T>::load(grad_in + i * V, gi);
#pragma unroll
                for (int j = 0; j < V; ++j) gv[j] += gi[j];
            }
            Vec<T>::store(grad_out + i * V, gv);
        }
    }
    const double part = block_sum(acc, red);
    if (threadIdx.x == 0) {
        atomicAdd(&ws->acc, part);
        __threadfence();
        if (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) {
            __threadfence();
            const double tot = atomicAdd(&ws->acc, 0.0);
            const float recon = static_cast<float>(weight * tot / static_cast<double>(n));
            if (loss_out != nullptr) *loss_out = recon;
            if (total_out != nullptr)
                *total_out = recon + (w_kl_ptr ? *w_kl_ptr : 0.f) * (kl ? *kl : 0.f) + w_adv * (adv ? *adv : 0.f);
            ws->acc = 0.0;          // re-armed for the next launch on this workspace
            ws->ticket = 0;
        }
    }
}

// total = recon + w_kl * kl + w_adv * adv   (vaegan_code.py:117), all scalars on the device
__global__ void total_loss_kernel(const float* recon, const float* kl, const float* adv, const float* w_kl_ptr,
                                  float w_kl, float w_adv, float* total) {
    pdl_enter();
    const float wk = w_kl_ptr ? *w_kl_ptr : w_kl;
    *total = *recon + wk * *kl + w_adv * *adv;
}

// ---- Adam (torch.optim.Adam defaults path: no amsgrad, no weight decay), one flat buffer per optimizer
__global__ void adam_tick_kernel(long long* step) { pdl_enter(); *step += 1; }

__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, long long n,
                                                       double lr_d, double b1_d, double b2_d, double eps_d,
                                                       const long long* __restrict__ step_ptr, float grad_scale) {
    pdl_enter();
    // hyper-parameters arrive as doubles and are narrowed exactly where torch narrows its Python floats
    const double t = static_cast<double>(*step_ptr);
    const float b1 = static_cast<float>(b1_d), b2 = static_cast<float>(b2_d), eps = static_cast<float>(eps_d);
    const float omb1 = static_cast<float>(1.0 - b1_d), omb2 = static_cast<float>(1.0 - b2_d);
    const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(b2_d, t)));
    const float step_size = static_cast<float>(lr_d / (1.0 - pow(b1_d, t)));
    const long long nvec = n / 4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i],
               vv = reinterpret_cast<float4*>(v)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float pp[4] = {pv.x, pv.y, pv.z, pv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
        const float gg[4] = {gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mm[j] = mm[j] + (gg[j] - mm[j]) * omb1;                // exp_avg.lerp_(grad, 1-beta1)
            vq[j] = vq[j] * b2 + omb2 * gg[j] * gg[j];             // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
            const float denom = sqrtf(vq[j]) / bc2_sqrt + eps;
            pp[j] = pp[j] - step_size * (mm[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * 4 + threadIdx.x; i < n; i += blockDim.x) {
            const float gi = g[i] * grad_scale;
            const float mi = m[i] + (gi - m[i]) * omb1;
            const float vi = v[i] * b2 + omb2 * gi * gi;
            m[i] = mi;
            v[i] = vi;
            p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        }
    }
}

// ---- Philox4x32-10 + Box-Muller standard normals
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
}
__global__ void __launch_bounds__(kThreads) randn_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                                        const unsigned long long* __restrict__ offset_ptr,
                                                        unsigned long long stream_id) {
    pdl_enter();
    const unsigned long long off = offset_ptr ? *offset_ptr : 0ull;
    const long long nquad = (n + 3) / 4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nquad;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint32_t c[4] = {static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(off),
                         static_cast<uint32_t>(off >> 32) ^ static_cast<uint32_t>(stream_id)};
        uint32_t k[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
#pragma unroll
        for (int r = 0; r < 10; ++r) philox_round(c, k);
        float z[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float u1 = (static_cast<float>(c[2 * h]) + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
            const float u2 = static_cast<float>(c[2 * h + 1]) * 2.3283064365386963e-10f;
            const float r = sqrtf(-2.f * logf(u1));
            float s, co;
            sincospif(2.f * u2, &s, &co);
            z[2 * h] = r * co;
            z[2 * h + 1] = r * s;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i * 4 + j < n) out[i * 4 + j] = z[j];
    }
}
__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { pdl_enter(); *c += inc; }

int grid_for(long long n) {
    return static_cast<int>(std::max<long long>(1, std::min<long long>(148 * 8, (n + kThreads - 1) / kThreads)));
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" int vg_reparam_fwd(const float* mu, const float* logvar, const float* eps, int batch, int nz, void* z,
                              VgDType z_dt, float* kl_out, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (mu == nullptr || logvar == nullptr || eps == nullptr || z == nullptr)
        return fail(VG_ERR_ARG, "reparam_fwd: null pointer");
    const int n = batch * nz;
    if (z_dt == VG_BF16)
        launch_k(reparam_fwd_kernel<__nv_bfloat16>, dim3(std::max(1, std::min(148, (n + 255) / 256))), dim3(256), 0, as_stream(stream), mu, logvar, eps, n, batch, static_cast<__nv_bfloat16*>(z), kl_out);
    else
        launch_k(reparam_fwd_kernel<float>, dim3(std::max(1, std::min(148, (n + 255) / 256))), dim3(256), 0, as_stream(stream), mu, logvar, eps, n, batch, static_cast<float*>(z), kl_out);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_reparam_bwd(const void* dz, VgDType dz_dt, const float* mu, const float* logvar, const float* eps,
                              int batch, int nz, const float* kl_weight_dev, float kl_weight, float* dmu,
                              float* dlogvar, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (dz == nullptr || mu == nullptr || logvar == nullptr || eps == nullptr || dmu == nullptr || dlogvar == nullptr)
        return fail(VG_ERR_ARG, "reparam_bwd: null pointer");
    const int n = batch * nz;
    if (dz_dt == VG_BF16)
        launch_k(reparam_bwd_kernel<__nv_bfloat16>, dim3(grid_for(n)), dim3(kThreads), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(dz), mu, logvar, eps, n, batch, kl_weight_dev, kl_weight, dmu, dlogvar);
    else
        launch_k(reparam_bwd_kernel<float>, dim3(grid_for(n)), dim3(kThreads), 0, as_stream(stream), static_cast<const float*>(dz), mu, logvar, eps, n, batch, kl_weight_dev, kl_weight, dmu, dlogvar);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_bce(const float* p, int n, float target, float weight, float* loss_out, int accumulate, float* dp,
                      void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (p == nullptr) return fail(VG_ERR_ARG, "bce: null pointer");
    launch_k(bce_kernel, dim3(1), dim3(1024), 0, as_stream(stream), p, n, target, weight, loss_out, accumulate, dp);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_bce_pair(const float* p, int n, float target_real, float target_fake, float weight, float* loss_out,
                           float* dp, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (p == nullptr) return fail(VG_ERR_ARG, "bce_pair: null pointer");
    launch_k(bce_pair_kernel, dim3(1), dim3(1024), 0, as_stream(stream), p, n, target_real, target_fake, weight, loss_out, dp);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_mse_total(const void* a, const void* b, VgDType dt, long long n, float weight, const void* grad_in,
                            void* grad_out, float* loss_out, const float* kl, const float* adv, const float* w_kl_dev,
                            float w_adv, float* total_out, void* ws, size_t ws_bytes, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (a == nullptr || b == nullptr) return fail(VG_ERR_ARG, "mse_total: null pointer");
    if (ws == nullptr || ws_bytes < sizeof(MseAcc)) return fail(VG_ERR_WORKSPACE, "mse_total: workspace too small");
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(grad_in) |
         reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(ws)) & 15)
        return fail(VG_ERR_ALIGN, "mse_total: 16-byte alignment");
    const int V = dt == VG_BF16 ? 8 : 4;
    if (n % V != 0) return fail(VG_ERR_SHAPE, "mse_total: element count must be a multiple of %d", V);
    const int blocks = grid_for(n / V);
    if (dt == VG_BF16)
        launch_k(mse_total_kernel<__nv_bfloat16>, dim3(blocks), dim3(kThreads), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), n, weight,
            static_cast<const __nv_bfloat16*>(grad_in), static_cast<__nv_bfloat16*>(grad_out), loss_out, kl, adv,
            w_kl_dev, w_adv, total_out, static_cast<MseAcc*>(ws));
    else
        launch_k(mse_total_kernel<float>, dim3(blocks), dim3(kThreads), 0, as_stream(stream), static_cast<const float*>(a), static_cast<const float*>(b), n, weight, static_cast<const float*>(grad_in),
            static_cast<float*>(grad_out), loss_out, kl, adv, w_kl_dev, w_adv, total_out, static_cast<MseAcc*>(ws));
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" size_t vg_mse_workspace_bytes(void) { return 148 * 8 * sizeof(double); }

extern "C" int vg_mse(const float* a, const float* b, long long n, float weight, const float* grad_in, float* grad_out,
                      float* loss_out, void* ws, size_t ws_bytes, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (a == nullptr || b == nullptr || loss_out == nullptr) return fail(VG_ERR_ARG, "mse: null pointer");
    if (ws == nullptr || ws_bytes < vg_mse_workspace_bytes()) return fail(VG_ERR_WORKSPACE, "mse: workspace too small");
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(grad_in) |
         reinterpret_cast<uintptr_t>(grad_out)) & 15)
        return fail(VG_ERR_ALIGN, "mse: 16-byte alignment");
    const int blocks = grid_for(n / 4 + 1);
    launch_k(mse_partial_kernel, dim3(blocks), dim3(kThreads), 0, as_stream(stream), a, b, n, weight, grad_in, grad_out,
                                                                   static_cast<double*>(ws));
    VG_LAUNCHED();
    launch_k(mse_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), static_cast<const double*>(ws), blocks, n, loss_out);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_total_loss(const float* recon, const float* kl, const float* adv, const float* w_kl_dev, float w_kl,
                             float w_adv, float* total, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    launch_k(total_loss_kernel, dim3(1), dim3(1), 0, as_stream(stream), recon, kl, adv, w_kl_dev, w_kl, w_adv, total);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                            double beta2, double eps, long long* step_dev, float grad_scale, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || step_dev == nullptr)
        return fail(VG_ERR_ARG, "adam: null pointer");
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
         reinterpret_cast<uintptr_t>(v)) & 15)
        return fail(VG_ERR_ALIGN, "adam: 16-byte alignment");
    launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_dev);
    VG_LAUNCHED();
    launch_k(adam_kernel, dim3(grid_for(n / 4 + 1)), dim3(kThreads), 0, as_stream(stream), p, g, m, v, n, lr, beta1, beta2, eps, step_dev,
                                                                         grad_scale);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_adam_apply(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                             double beta2, double eps, const long long* step_dev, float grad_scale, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || step_dev == nullptr)
        return fail(VG_ERR_ARG, "adam: null pointer");
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
         reinterpret_cast<uintptr_t>(v)) & 15)
        return fail(VG_ERR_ALIGN, "adam: 16-byte alignment");
    if (n <= 0) return VG_OK;
    launch_k(adam_kernel, dim3(grid_for(n / 4 + 1)), dim3(kThreads), 0, as_stream(stream), p, g, m, v, n, lr, beta1, beta2,
             eps, step_dev, grad_scale);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_adam_tick(long long* step_dev, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (step_dev == nullptr) return fail(VG_ERR_ARG, "adam_tick: null pointer");
    launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_dev);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_randn(float* out, long long n, unsigned long long seed, unsigned long long* offset_dev,
                        unsigned long long stream_id, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (out == nullptr) return fail(VG_ERR_ARG, "randn: null pointer");
    launch_k(randn_kernel, dim3(grid_for((n + 3) / 4)), dim3(kThreads), 0, as_stream(stream), out, n, seed, offset_dev, stream_id);
    VG_LAUNCHED();
    if (offset_dev != nullptr) {
        launch_k(counter_add_kernel, dim3(1), dim3(1), 0, as_stream(stream), offset_dev, 1ull);
        VG_LAUNCHED();
    }
    return VG_OK;
}
