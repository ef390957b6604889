// Native self-check of the convolution C-ABI (include/vaegan_b200.h) against straightforward CPU loops.
// Build:  make -C tests/native      Run on a B200:  build/igemm_harness [perf]
// Exit code 0 only if every case passes.  This is test infrastructure, not product code.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vaegan_b200.h"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

static uint32_t rng_state = 12345u;
static float frand() {  // uniform in [-1, 1)
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct Case {
    const char* name;
    VgConvGeom g;
};

static VgConvGeom geom(int B, int bh, int bw, int bc, int sc, int k, int s, int p) {
    VgConvGeom g;
    g.batch = B;
    g.big_h = bh;
    g.big_w = bw;
    g.big_c = bc;
    g.small_h = (bh + 2 * p - k) / s + 1;
    g.small_w = (bw + 2 * p - k) / s + 1;
    g.small_c = sc;
    g.kernel = k;
    g.stride = s;
    g.pad = p;
    g.big_c_valid = 0;
    return g;
}

template <typename T>
static T* to_dev(const std::vector<T>& h) {
    T* d;
    CK(cudaMalloc(&d, h.size() * sizeof(T) + 256));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
    std::vector<__nv_bfloat16> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
    return o;
}

static bool report(const char* what, const char* name, double max_err, double max_ref, double tol) {
    const bool ok = max_err <= tol * (max_ref > 1.0 ? max_ref : 1.0) && std::isfinite(max_err);
    printf("  %-6s %-34s max|err| %.3e  max|ref| %.3e  %s\n", what, name, max_err, max_ref, ok ? "PASS" : "FAIL");
    return ok;
}

// Run all three contractions for one geometry in both arithmetic modes.
static bool run_case(const Case& c, bool verbose_fail) {
    const VgConvGeom& g = c.g;
    const int kk = g.kernel * g.kernel;
    const size_t n_big = (size_t)g.batch * g.big_h * g.big_w * g.big_c;
    const size_t n_small = (size_t)g.batch * g.small_h * g.small_w * g.small_c;
    const size_t n_w = (size_t)g.small_c * g.big_c * kk;
    std::vector<float> big(n_big), small(n_small), w(n_w), bias(g.small_c);
    for (auto& v : big) v = bf16_round(frand());
    for (auto& v : small) v = bf16_round(frand());
    for (auto& v : w) v = bf16_round(frand() * 0.25f);
    for (auto& v : bias) v = frand();

    // CPU references (double accumulation over the bf16-representable inputs)
    std::vector<double> r_down(n_small, 0.0), r_up(n_big, 0.0), r_wg(n_w, 0.0);
    for (int b = 0; b < g.batch; ++b)
        for (int oy = 0; oy < g.small_h; ++oy)
            for (int ox = 0; ox < g.small_w; ++ox)
                for (int ky = 0; ky < g.kernel; ++ky)
                    for (int kx = 0; kx < g.kernel; ++kx) {
                        const int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
                        if (iy < 0 || iy >= g.big_h || ix < 0 || ix >= g.big_w) continue;
                        const float* bp = &big[(((size_t)b * g.big_h + iy) * g.big_w + ix) * g.big_c];
                        const size_t so = (((size_t)b * g.small_h + oy) * g.small_w + ox) * g.small_c;
                        double* up = &r_up[(((size_t)b * g.big_h + iy) * g.big_w + ix) * g.big_c];
                        for (int sc = 0; sc < g.small_c; ++sc) {
                            const float sv = small[so + sc];
                            double acc = 0.0;
                            for (int bc = 0; bc < g.big_c; ++bc) {
                                const size_t wi = ((size_t)sc * g.big_c + bc) * kk + ky * g.kernel + kx;
                                acc += (double)bp[bc] * w[wi];
                                up[bc] += (double)sv * w[wi];
                                r_wg[wi] += (double)sv * bp[bc];
                            }
                            r_down[so + sc] += acc;
                        }
                    }
    for (size_t i = 0; i < n_small; ++i) r_down[i] += bias[i % g.small_c];

    bool ok = true;
    for (int mode = 0; mode < 2; ++mode) {
        const VgDType dt = mode == 0 ? VG_F32 : VG_BF16;
        const size_t es = mode == 0 ? 4 : 2;
        void *d_big, *d_small, *d_out_small, *d_out_big, *d_wd = nullptr, *d_wu = nullptr;
        float *d_w = to_dev(w), *d_bias = to_dev(bias), *d_dw;
        if (mode == 0) {
            d_big = to_dev(big);
            d_small = to_dev(small);
        } else {
            d_big = to_dev(to_bf16(big));
            d_small = to_dev(to_bf16(small));
            CK(cudaMalloc(&d_wd, n_w * 2));
            CK(cudaMalloc(&d_wu, n_w * 2));
            int rc = vg_pack_weights_bf16(&g, d_w, d_wd, d_wu, nullptr);
            if (rc != VG_OK) { printf("  pack failed: %s\n", vg_last_error()); return false; }
        }
        CK(cudaMalloc(&d_out_small, n_small * 4));
        CK(cudaMalloc(&d_out_big, n_big * 4));
        CK(cudaMalloc(&d_dw, n_w * 4));
        void* d_ws = nullptr;
        const size_t ws_bytes = vg_conv_down_workspace_bytes(&g);
        CK(cudaMalloc(&d_ws, ws_bytes));
        void* d_wws = nullptr;
        const size_t wws_bytes = vg_conv_wgrad_workspace_bytes(&g, dt);
        CK(cudaMalloc(&d_wws, wws_bytes + 16));
        CK(cudaMemset(d_out_small, 0xFF, n_small * 4));
        CK(cudaMemset(d_out_big, 0xFF, n_big * 4));
        CK(cudaMemset(d_dw, 0, n_w * 4));
        const void* w_down = mode == 0 ? (const void*)d_w : d_wd;
        const void* w_up = mode == 0 ? (const void*)d_w : d_wu;
        int rc;
        rc = vg_conv_down(&g, dt, d_big, w_down, d_bias, d_out_small, 0, d_ws, ws_bytes, nullptr);
        if (rc != VG_OK) { printf("  down failed (%d): %s\n", rc, vg_last_error()); ok = false; }
        rc = vg_conv_up(&g, dt, d_small, w_up, d_out_big, nullptr);
        if (rc != VG_OK) { printf("  up failed (%d): %s\n", rc, vg_last_error()); ok = false; }
        rc = vg_conv_wgrad(&g, dt, d_small, d_big, d_dw, d_wws, wws_bytes, nullptr);
        if (rc != VG_OK) { printf("  wgrad failed (%d): %s\n", rc, vg_last_error()); ok = false; }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("  kernel fault in case %s mode %d: %s\n", c.name, mode, cudaGetErrorString(e));
            exit(3);
        }
        std::vector<uint8_t> h_small(n_small * es), h_big(n_big * es);
        std::vector<float> h_dw(n_w);
        CK(cudaMemcpy(h_small.data(), d_out_small, n_small * es, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_big.data(), d_out_big, n_big * es, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_dw.data(), d_dw, n_w * 4, cudaMemcpyDeviceToHost));
        auto get = [&](const std::vector<uint8_t>& v, size_t i) -> double {
            if (mode == 0) return reinterpret_cast<const float*>(v.data())[i];
            return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(v.data())[i]);
        };
        double e_d = 0, m_d = 0, e_u = 0, m_u = 0, e_w = 0, m_w = 0;
        size_t bad_d = 0, bad_u = 0, bad_w = 0;
        for (size_t i = 0; i < n_small; ++i) {
            const double err = fabs(get(h_small, i) - r_down[i]);
            if (!(err <= e_d)) { e_d = err; bad_d = i; }
            m_d = fmax(m_d, fabs(r_down[i]));
        }
        for (size_t i = 0; i < n_big; ++i) {
            const double err = fabs(get(h_big, i) - r_up[i]);
            if (!(err <= e_u)) { e_u = err; bad_u = i; }
            m_u = fmax(m_u, fabs(r_up[i]));
        }
        for (size_t i = 0; i < n_w; ++i) {
            const double err = fabs(h_dw[i] - r_wg[i]);
            if (!(err <= e_w)) { e_w = err; bad_w = i; }
            m_w = fmax(m_w, fabs(r_wg[i]));
        }
        // fp32 mode: accumulation-order noise only; bf16 mode: output rounding (2^-9 relative) for down/up
        const double tol_out = mode == 0 ? 2e-5 : 6e-3, tol_w = 1e-4;
        std::string nm = std::string(c.name) + (mode == 0 ? " [f32]" : " [bf16]");
        const bool o1 = report("down", nm.c_str(), e_d, m_d, tol_out);
        const bool o2 = report("up", nm.c_str(), e_u, m_u, tol_out);
        const bool o3 = report("wgrad", nm.c_str(), e_w, m_w, tol_w);
        // VG_WGRAD_OVERWRITE: the destination starts as garbage (NaN pattern) and must come back as the plain gradient
        CK(cudaMemset(d_dw, 0xFF, n_w * 4));
        rc = vg_conv_wgrad_ex(&g, dt, d_small, d_big, d_dw, d_wws, wws_bytes, VG_WGRAD_OVERWRITE, nullptr);
        if (rc != VG_OK) { printf("  wgrad(overwrite) failed (%d): %s\n", rc, vg_last_error()); ok = false; }
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h_dw.data(), d_dw, n_w * 4, cudaMemcpyDeviceToHost));
        double e_o = 0;
        for (size_t i = 0; i < n_w; ++i) {
            const double err = fabs(h_dw[i] - r_wg[i]);
            if (!(err <= e_o)) e_o = err;
        }
        const bool o4 = report("wgrad=", nm.c_str(), e_o, m_w, tol_w);
        ok = ok && o4;
        if (verbose_fail) {
            if (!o1) printf("    down worst at %zu: got %g want %g\n", bad_d, get(h_small, bad_d), r_down[bad_d]);
            if (!o2) printf("    up worst at %zu: got %g want %g\n", bad_u, get(h_big, bad_u), r_up[bad_u]);
            if (!o3) printf("    wgrad worst at %zu: got %g want %g\n", bad_w, (double)h_dw[bad_w], r_wg[bad_w]);
        }
        ok = ok && o1 && o2 && o3;
        cudaFree(d_big); cudaFree(d_small); cudaFree(d_out_small); cudaFree(d_out_big); cudaFree(d_w);
        cudaFree(d_bias); cudaFree(d_dw); cudaFree(d_ws); cudaFree(d_wws);
        if (d_wd) cudaFree(d_wd);
        if (d_wu) cudaFree(d_wu);
    }
    return ok;
}

#ifdef VG_WGRAD_TRACE
extern "C" int vg_debug_wgrad_trace(unsigned long long* out8, int reset);
#endif
static void perf_case(const char* name, VgConvGeom g, int iters) {
    const int kk = g.kernel * g.kernel;
    const size_t n_big = (size_t)g.batch * g.big_h * g.big_w * g.big_c;
    const size_t n_small = (size_t)g.batch * g.small_h * g.small_w * g.small_c;
    const size_t n_w = (size_t)g.small_c * g.big_c * kk;
    void *d_big, *d_small, *d_wd, *d_wu;
    float *d_w, *d_dw;
    CK(cudaMalloc(&d_big, n_big * 2));
    CK(cudaMalloc(&d_small, n_small * 2));
    CK(cudaMalloc(&d_wd, n_w * 2));
    CK(cudaMalloc(&d_wu, n_w * 2));
    CK(cudaMalloc(&d_w, n_w * 4));
    CK(cudaMalloc(&d_dw, n_w * 4));
    CK(cudaMemset(d_big, 0, n_big * 2));
    CK(cudaMemset(d_small, 0, n_small * 2));
    CK(cudaMemset(d_w, 0, n_w * 4));
    CK(cudaMemset(d_dw, 0, n_w * 4));
    vg_pack_weights_bf16(&g, d_w, d_wd, d_wu, nullptr);
    void* d_wws = nullptr;
    const size_t wws_bytes = vg_conv_wgrad_workspace_bytes(&g, VG_BF16);
    CK(cudaMalloc(&d_wws, wws_bytes + 16));
    const double flops = 2.0 * g.batch * g.small_h * g.small_w * (double)g.small_c * g.big_c * kk;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int op = 0; op < 3; ++op) {
        auto run = [&]() {
            if (op == 0) return vg_conv_down(&g, VG_BF16, d_big, d_wd, nullptr, d_small, 0, nullptr, 0, nullptr);
            if (op == 1) return vg_conv_up(&g, VG_BF16, d_small, d_wu, d_big, nullptr);
            return vg_conv_wgrad(&g, VG_BF16, d_small, d_big, d_dw, d_wws, wws_bytes, nullptr);
        };
        for (int i = 0; i < 3; ++i)
            if (run() != VG_OK) { printf("perf %s failed: %s\n", name, vg_last_error()); return; }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) run();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
        printf("  perf %-28s %-5s %8.1f us  %7.1f TFLOP/s\n", name, op == 0 ? "down" : (op == 1 ? "up" : "wgrad"),
               ms * 1e3, flops / (ms * 1e-3) / 1e12);
#ifdef VG_WGRAD_TRACE
        if (op == 2) {
            unsigned long long t[8];
            vg_debug_wgrad_trace(t, 1);
            const double n = iters + 3;
            printf("    trace (cycles per launch, CTA 0): producer wait_empty_b %.0f  q_tma %.0f  wait_empty_a %.0f  p_tma %.0f"
                   " | issuer wait_full_b %.0f  umma %.0f  commit %.0f  wait_full_a %.0f\n",
                   t[0] / n, t[1] / n, t[2] / n, t[3] / n, t[4] / n, t[5] / n, t[6] / n, t[7] / n);
        }
#endif
    }
    cudaFree(d_big); cudaFree(d_small); cudaFree(d_wd); cudaFree(d_wu); cudaFree(d_w); cudaFree(d_dw); cudaFree(d_wws);
}

// ---- timing of the fused-epilogue launches (modes 1-3) and of the weight gradient on RANDOM data (zero operands draw
// less power and clock higher than a real step does)
__global__ void fill_random_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = __float2bfloat16_rn(((h & 0xFFFF) / 32768.0f - 1.0f) * scale);
    }
}
__global__ void fill_f32(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static int g_fused_only = -1, g_fused_idx = 0;      // `fused <n>`: run only the n-th fused case (for ncu)
static void perf_fused(const char* name, VgConvGeom g, int up, int mode, int groups, int iters) {
    if (g_fused_only >= 0 && g_fused_idx++ != g_fused_only) return;
    const int kk = g.kernel * g.kernel;
    const size_t n_big = (size_t)g.batch * g.big_h * g.big_w * g.big_c;
    const size_t n_small = (size_t)g.batch * g.small_h * g.small_w * g.small_c;
    const size_t n_w = (size_t)g.small_c * g.big_c * kk;
    const size_t n_out = up ? n_big : n_small;
    const int C = up ? g.big_c : g.small_c;
    __nv_bfloat16 *d_big, *d_small, *d_wd, *d_wu, *d_x;
    float *d_w, *d_sums, *d_stats, *d_dw;
    CK(cudaMalloc(&d_big, n_big * 2));
    CK(cudaMalloc(&d_small, n_small * 2));
    CK(cudaMalloc(&d_x, n_out * 2));
    CK(cudaMalloc(&d_wd, n_w * 2));
    CK(cudaMalloc(&d_wu, n_w * 2));
    CK(cudaMalloc(&d_w, n_w * 4));
    CK(cudaMalloc(&d_dw, n_w * 4));
    CK(cudaMalloc(&d_sums, (size_t)groups * 2 * C * 4));
    CK(cudaMalloc(&d_stats, (size_t)groups * 4 * C * 4));
    fill_random_bf16<<<1024, 256>>>(d_big, n_big, 1u, 1.f);
    fill_random_bf16<<<1024, 256>>>(d_small, n_small, 2u, 1.f);
    fill_random_bf16<<<1024, 256>>>(d_x, n_out, 3u, 1.f);
    fill_random_bf16<<<1024, 256>>>(d_wd, n_w, 4u, 0.05f);
    fill_random_bf16<<<1024, 256>>>(d_wu, n_w, 5u, 0.05f);
    fill_f32<<<64, 256>>>(d_stats, (size_t)groups * 4 * C, 0.5f);
    CK(cudaMemset(d_sums, 0, (size_t)groups * 2 * C * 4));
    CK(cudaMemset(d_dw, 0, n_w * 4));
    VgEpilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.mode = mode;
    ep.groups = groups;
    ep.channels = C;
    ep.act = VG_ACT_LEAKY;
    ep.slope = 0.2f;
    ep.sums = d_sums;
    ep.x = d_x;
    ep.stats = d_stats;
    const double flops = 2.0 * g.batch * g.small_h * g.small_w * (double)g.small_c * g.big_c * kk;
    const double bytes = 2.0 * (n_big + n_small + n_w + (mode >= 2 ? n_out : 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&]() {
        if (mode == 0) return up ? vg_conv_up(&g, VG_BF16, d_small, d_wu, d_big, nullptr)
                                 : vg_conv_down(&g, VG_BF16, d_big, d_wd, nullptr, d_small, 0, nullptr, 0, nullptr);
        return up ? vg_conv_up_ex(&g, VG_BF16, d_small, d_wu, d_big, &ep, nullptr)
                  : vg_conv_down_ex(&g, VG_BF16, d_big, d_wd, nullptr, d_small, &ep, nullptr);
    };
    for (int i = 0; i < 3; ++i)
        if (run() != VG_OK) { printf("  fused %s failed: %s\n", name, vg_last_error()); return; }
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) run();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
        best = ms < best ? ms : best;
        sum += ms;
    }
    printf("  fused %-30s %-4s mode %d  %8.1f us (mean %.1f)  %7.1f TFLOP/s  %6.0f GB/s algorithmic\n", name,
           up ? "up" : "down", mode, best * 1e3, sum / 3 * 1e3, flops / (best * 1e-3) / 1e12, bytes / (best * 1e-3) / 1e9);
    cudaFree(d_big); cudaFree(d_small); cudaFree(d_x); cudaFree(d_wd); cudaFree(d_wu); cudaFree(d_w); cudaFree(d_dw);
    cudaFree(d_sums); cudaFree(d_stats);
}

static void perf_wgrad_random(const char* name, VgConvGeom g, int iters) {
    const int kk = g.kernel * g.kernel;
    const size_t n_big = (size_t)g.batch * g.big_h * g.big_w * g.big_c;
    const size_t n_small = (size_t)g.batch * g.small_h * g.small_w * g.small_c;
    const size_t n_w = (size_t)g.small_c * g.big_c * kk;
    __nv_bfloat16 *d_big, *d_small;
    float* d_dw;
    CK(cudaMalloc(&d_big, n_big * 2));
    CK(cudaMalloc(&d_small, n_small * 2));
    CK(cudaMalloc(&d_dw, n_w * 4));
    fill_random_bf16<<<1024, 256>>>(d_big, n_big, 1u, 1.f);
    fill_random_bf16<<<1024, 256>>>(d_small, n_small, 2u, 1.f);
    void* d_wws = nullptr;
    const size_t wws_bytes = vg_conv_wgrad_workspace_bytes(&g, VG_BF16);
    CK(cudaMalloc(&d_wws, wws_bytes + 16));
    const double flops = 2.0 * g.batch * g.small_h * g.small_w * (double)g.small_c * g.big_c * kk;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int flags = 0; flags < 2; ++flags) {
        auto run = [&]() { return vg_conv_wgrad_ex(&g, VG_BF16, d_small, d_big, d_dw, d_wws, wws_bytes, flags, nullptr); };
        CK(cudaMemset(d_dw, 0, n_w * 4));
        for (int i = 0; i < 3; ++i)
            if (run() != VG_OK) { printf("  wgrad %s failed: %s\n", name, vg_last_error()); return; }
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            for (int i = 0; i < iters; ++i) run();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ms /= iters;
            best = ms < best ? ms : best;
        }
        printf("  wgrad %-30s %-9s %8.1f us  %7.1f TFLOP/s  (scratch %zu KB)\n", name, flags ? "overwrite" : "add",
               best * 1e3, flops / (best * 1e-3) / 1e12, wws_bytes >> 10);
    }
    cudaFree(d_big); cudaFree(d_small); cudaFree(d_dw); cudaFree(d_wws);
}

int main(int argc, char** argv) {
    const bool perf = argc > 1 && strcmp(argv[1], "perf") == 0;
    const bool fused = argc > 1 && strcmp(argv[1], "fused") == 0;
    if (fused) {
        if (argc > 2) g_fused_only = atoi(argv[2]);
        if (vg_device_check() != VG_OK) { printf("device check failed: %s\n", vg_last_error()); return 1; }
        printf("fused-epilogue launches, random data (cfg-2 shapes; D layers at the stacked batch 512)\n");
        perf_fused("G last dgrad (s2d 64->64)", geom(256, 64, 64, 64, 64, 4, 2, 1), 1, 2, 1, 20);
        perf_fused("G last dgrad (s2d 64->64)", geom(256, 64, 64, 64, 64, 4, 2, 1), 1, 0, 1, 20);
        perf_fused("D 64->128 dgrad B512", geom(512, 32, 32, 64, 128, 4, 2, 1), 1, 3, 1, 20);
        perf_fused("D 64->128 dgrad B512", geom(512, 32, 32, 64, 128, 4, 2, 1), 1, 0, 1, 20);
        perf_fused("D 128->256 dgrad B512", geom(512, 16, 16, 128, 256, 4, 2, 1), 1, 2, 2, 20);
        perf_fused("D 128->256 dgrad B512", geom(512, 16, 16, 128, 256, 4, 2, 1), 1, 0, 2, 20);
        perf_fused("D 256->512 dgrad B512", geom(512, 8, 8, 256, 512, 4, 2, 1), 1, 2, 2, 20);
        perf_fused("G 128->64 dgrad", geom(256, 64, 64, 64, 128, 4, 2, 1), 0, 2, 1, 20);
        perf_fused("G 128->64 dgrad", geom(256, 64, 64, 64, 128, 4, 2, 1), 0, 0, 1, 20);
        perf_fused("G 256->128 dgrad", geom(256, 32, 32, 128, 256, 4, 2, 1), 0, 2, 1, 20);
        perf_fused("G 256->128 dgrad", geom(256, 32, 32, 128, 256, 4, 2, 1), 0, 0, 1, 20);
        perf_fused("G 512->256 dgrad", geom(256, 16, 16, 256, 512, 4, 2, 1), 0, 2, 1, 20);
        perf_fused("G 1024->512 dgrad", geom(256, 8, 8, 512, 1024, 4, 2, 1), 0, 2, 1, 20);
        perf_fused("G 128->64 fwd", geom(256, 64, 64, 64, 128, 4, 2, 1), 1, 1, 1, 20);
        perf_fused("G 128->64 fwd", geom(256, 64, 64, 64, 128, 4, 2, 1), 1, 0, 1, 20);
        perf_fused("G 256->128 fwd", geom(256, 32, 32, 128, 256, 4, 2, 1), 1, 1, 1, 20);
        perf_fused("G 512->256 fwd", geom(256, 16, 16, 256, 512, 4, 2, 1), 1, 1, 1, 20);
        perf_fused("G 1024->512 fwd", geom(256, 8, 8, 512, 1024, 4, 2, 1), 1, 1, 1, 20);
        perf_fused("D 64->128 fwd B512", geom(512, 32, 32, 64, 128, 4, 2, 1), 0, 1, 2, 20);
        perf_fused("D 128->256 fwd B512", geom(512, 16, 16, 128, 256, 4, 2, 1), 0, 1, 2, 20);
        perf_fused("D 256->512 fwd B512", geom(512, 8, 8, 256, 512, 4, 2, 1), 0, 1, 2, 20);
        if (g_fused_only >= 0) return 0;
        printf("weight gradients, random data\n");
        perf_wgrad_random("G 1024->512", geom(256, 8, 8, 512, 1024, 4, 2, 1), 20);
        perf_wgrad_random("G 512->256", geom(256, 16, 16, 256, 512, 4, 2, 1), 20);
        perf_wgrad_random("G 256->128", geom(256, 32, 32, 128, 256, 4, 2, 1), 20);
        perf_wgrad_random("G 128->64", geom(256, 64, 64, 64, 128, 4, 2, 1), 20);
        perf_wgrad_random("G last (s2d 64->64)", geom(256, 64, 64, 64, 64, 4, 2, 1), 20);
        perf_wgrad_random("D 64->128 B512", geom(512, 32, 32, 64, 128, 4, 2, 1), 20);
        perf_wgrad_random("D 128->256 B512", geom(512, 16, 16, 128, 256, 4, 2, 1), 20);
        perf_wgrad_random("D 256->512 B512", geom(512, 8, 8, 256, 512, 4, 2, 1), 20);
        perf_wgrad_random("D first (s2d 64->64 k2) B512", geom(512, 33, 33, 64, 64, 2, 1, 0), 20);
        return 0;
    }
    int rc = vg_device_check();
    if (rc != VG_OK) {
        printf("device check failed: %s\n", vg_last_error());
        return 1;
    }
    std::vector<Case> cases = {
        {"k4s2p1 16x16 c64->128 B4", geom(4, 16, 16, 64, 128, 4, 2, 1)},
        {"k4s2p1 8x8 c128->256 B8", geom(8, 8, 8, 128, 256, 4, 2, 1)},
        {"k4s2p0 31x31 c32->64 B3 (odd)", geom(3, 31, 31, 32, 64, 4, 2, 0)},
        {"k4s2p0 14x14 c64->128 B5 (odd)", geom(5, 14, 14, 64, 128, 4, 2, 0)},
        {"k4s2p0 6x6 c128->256 B9", geom(9, 6, 6, 128, 256, 4, 2, 0)},
        {"k4s1p0 4x4 c256<-128 dense B16", geom(16, 4, 4, 256, 128, 4, 1, 0)},
        {"k2s1p0 2x2 c256->256 (fc) B12", geom(12, 2, 2, 256, 256, 2, 1, 0)},
        {"k4s2p1 32x32 c64->128 B2", geom(2, 32, 32, 64, 128, 4, 2, 1)},
        {"k4s2p1 4x4 c256->512 B8", geom(8, 4, 4, 256, 512, 4, 2, 1)},
        {"k4s2p1 64x64 c3->64 B2 (image)", geom(2, 64, 64, 3, 64, 4, 2, 1)},
        {"k3s1p1 16x16 c3<-64 B2 (image)", geom(2, 16, 16, 3, 64, 3, 1, 1)},
        {"k4s1p0 4x4 c512->1 B8 (gemv)", geom(8, 4, 4, 512, 1, 4, 1, 0)},
        {"k4s2p1 16x16 c16->32 B4", geom(4, 16, 16, 16, 32, 4, 2, 1)},
        {"k4s1p0 4x4 c1024->128 B64 (splitK)", geom(64, 4, 4, 1024, 128, 4, 1, 0)},
        {"k4s2p1 64x64 c16->64 B2 (padded)", geom(2, 64, 64, 16, 64, 4, 2, 1)},
        {"k3s1p1 32x32 c16<-64 B3 (padded)", geom(3, 32, 32, 16, 64, 3, 1, 1)},
        {"k4s2p1 16x16 c32->64 B4", geom(4, 16, 16, 32, 64, 4, 2, 1)},
        // shapes the halo-tile path accepts (VG_HALO=1: 128-byte rows, output grid a multiple of 16 x 8)
        {"k4s2p1 64x64 c64->64 B2 (halo)", geom(2, 64, 64, 64, 64, 4, 2, 1)},
        {"k4s2p1 32x48 c128->64 B3 (halo, 2 chunks)", geom(3, 32, 48, 128, 64, 4, 2, 1)},
        {"k3s1p1 32x16 c64<-64 B2 (halo 3x3)", geom(2, 32, 16, 64, 64, 3, 1, 1)},
    };
    bool all = true;
    for (const Case& c : cases) {
        printf("case %s\n", c.name);
        all = run_case(c, true) && all;
        fflush(stdout);
    }
    if (perf) {
        printf("perf (bf16, cfg-2 layer shapes, B=256)\n");
        perf_case("G 1024->512 4^2->8^2", geom(256, 8, 8, 512, 1024, 4, 2, 1), 20);
        perf_case("G 512->256 8^2->16^2", geom(256, 16, 16, 256, 512, 4, 2, 1), 20);
        perf_case("G 256->128 16^2->32^2", geom(256, 32, 32, 128, 256, 4, 2, 1), 20);
        perf_case("G 128->64 32^2->64^2", geom(256, 64, 64, 64, 128, 4, 2, 1), 20);
        perf_case("D 16pad->64 64^2->32^2 B512", geom(512, 64, 64, 16, 64, 4, 2, 1), 20);
        perf_case("G 64->16pad k3 64^2", geom(256, 64, 64, 16, 64, 3, 1, 1), 20);
        perf_case("D 256->512 8^2->4^2 B512", geom(512, 8, 8, 256, 512, 4, 2, 1), 20);
    }
    printf(all ? "ALL PASS\n" : "SOME FAILED\n");
    return all ? 0 : 1;
}
