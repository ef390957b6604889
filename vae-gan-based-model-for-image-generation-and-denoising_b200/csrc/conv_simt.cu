// CUDA-core implicit-GEMM convolution contractions (see conv_simt.cuh).
// One 64x64x16 smem-tiled GEMM skeleton; the three contractions differ only in how A(m,k), B(k,n) are gathered
// and where C(m,n) is stored.  fp32 accumulation; activations fp32 or bf16 (NHWC).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "pdl.cuh"
#include "conv_simt.cuh"

namespace vg {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(p)));
}
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct Geo {
    int B, bh, bw, bc, sh, sw, sc, k, s, p;
    int bcv;  // big-side channels present in the weight tensor (bc may be padded)
};

// ---- down: small[m=(b,oy,ox)][n=sc] = sum_{k=(tap,bc)} big(...) * w(sc, bc, tap)
template <typename T, typename WT>
struct DownOp {
    static constexpr bool kAContigK = true, kBContigK = true;
    Geo g;
    const T* big;
    const WT* w;
    const float* bias;
    void* out;
    int out_f32;
    long long ws_sc, ws_bc, ws_tap;  // weight element strides
    __device__ int M() const { return g.B * g.sh * g.sw; }
    __device__ int N() const { return g.sc; }
    __device__ void krange(int, int* k0, int* k1) const {
        *k0 = 0;
        *k1 = g.k * g.k * g.bc;
    }
    __device__ float a(int m, int kk, int) const {
        const int tap = kk / g.bc, c = kk - tap * g.bc;
        const int ky = tap / g.k, kx = tap - ky * g.k;
        const int ox = m % g.sw, t = m / g.sw, oy = t % g.sh, b = t / g.sh;
        const int iy = oy * g.s - g.p + ky, ix = ox * g.s - g.p + kx;
        if (iy < 0 || iy >= g.bh || ix < 0 || ix >= g.bw) return 0.f;
        return ldf(big + ((static_cast<long long>(b) * g.bh + iy) * g.bw + ix) * g.bc + c);
    }
    __device__ float b(int kk, int n, int) const {
        const int tap = kk / g.bc, c = kk - tap * g.bc;
        if (c >= g.bcv) return 0.f;
        return ldf(w + n * ws_sc + c * ws_bc + tap * ws_tap);
    }
    __device__ void store(int m, int n, float v, int) const {
        if (bias != nullptr) v += __ldg(bias + n);
        const long long o = static_cast<long long>(m) * g.sc + n;
        if (out_f32) static_cast<float*>(out)[o] = v;
        else stf(static_cast<T*>(out) + o, v);
    }
};

// ---- up: one launch slice per output parity phase z=(ay,ax); rows are the phase's pixels
template <typename T, typename WT>
struct UpOp {
    static constexpr bool kAContigK = true, kBContigK = true;
    Geo g;
    const T* small;
    const WT* w;
    T* out;
    long long ws_sc, ws_bc, ws_tap;
    __device__ int gh() const { return (g.bh + g.s - 1) / g.s; }
    __device__ int gw() const { return (g.bw + g.s - 1) / g.s; }
    __device__ int M() const { return g.B * gh() * gw(); }
    __device__ int N() const { return g.bc; }
    __device__ void phase(int z, int* ay, int* ax, int* ky0, int* kx0, int* nky, int* nkx) const {
        *ay = z / g.s;
        *ax = z % g.s;
        *ky0 = (*ay + g.p) % g.s;
        *kx0 = (*ax + g.p) % g.s;
        *nky = *ky0 < g.k ? (g.k - *ky0 + g.s - 1) / g.s : 0;
        *nkx = *kx0 < g.k ? (g.k - *kx0 + g.s - 1) / g.s : 0;
    }
    __device__ void krange(int z, int* k0, int* k1) const {
        int ay, ax, ky0, kx0, nky, nkx;
        phase(z, &ay, &ax, &ky0, &kx0, &nky, &nkx);
        *k0 = 0;
        *k1 = nky * nkx * g.sc;
    }
    __device__ float a(int m, int kk, int z) const {
        int ay, ax, ky0, kx0, nky, nkx;
        phase(z, &ay, &ax, &ky0, &kx0, &nky, &nkx);
        const int t = kk / g.sc, c = kk - t * g.sc;
        const int ky = ky0 + (t / nkx) * g.s, kx = kx0 + (t % nkx) * g.s;
        const int j = m % gw(), r = m / gw(), i = r % gh(), b = r / gh();
        const int sy = i + (ay + g.p - ky) / g.s, sx = j + (ax + g.p - kx) / g.s;  // exact divisions
        if (sy < 0 || sy >= g.sh || sx < 0 || sx >= g.sw) return 0.f;
        return ldf(small + ((static_cast<long long>(b) * g.sh + sy) * g.sw + sx) * g.sc + c);
    }
    __device__ float b(int kk, int n, int z) const {
        int ay, ax, ky0, kx0, nky, nkx;
        phase(z, &ay, &ax, &ky0, &kx0, &nky, &nkx);
        const int t = kk / g.sc, c = kk - t * g.sc;
        const int ky = ky0 + (t / nkx) * g.s, kx = kx0 + (t % nkx) * g.s;
        if (n >= g.bcv) return 0.f;
        return ldf(w + c * ws_sc + n * ws_bc + (ky * g.k + kx) * ws_tap);
    }
    __device__ void store(int m, int n, float v, int z) const {
        const int ay = z / g.s, ax = z % g.s;
        const int j = m % gw(), r = m / gw(), i = r % gh(), b = r / gh();
        const int y = i * g.s + ay, x = j * g.s + ax;
        if (y >= g.bh || x >= g.bw) return;
        stf(out + ((static_cast<long long>(b) * g.bh + y) * g.bw + x) * g.bc + n, v);
    }
};

// ---- wgrad: slice z = tap * splits + split; reduction over a slice of the small-side pixels
template <typename T>
struct WgradOp {
    static constexpr bool kAContigK = false, kBContigK = false;
    Geo g;
    const T* small;
    const T* big;
    float* dw;
    int splits;
    __device__ int M() const { return g.sc; }
    __device__ int N() const { return g.bc; }
    __device__ void krange(int z, int* k0, int* k1) const {
        const long long P = static_cast<long long>(g.B) * g.sh * g.sw;
        const int sp = z % splits;
        *k0 = static_cast<int>(P * sp / splits);
        *k1 = static_cast<int>(P * (sp + 1) / splits);
    }
    __device__ float a(int m, int pix, int) const { return ldf(small + static_cast<long long>(pix) * g.sc + m); }
    __device__ float b(int pix, int n, int z) const {
        const int tap = z / splits;
        const int ky = tap / g.k, kx = tap - ky * g.k;
        const int ox = pix % g.sw, t = pix / g.sw, oy = t % g.sh, b = t / g.sh;
        const int iy = oy * g.s - g.p + ky, ix = ox * g.s - g.p + kx;
        if (iy < 0 || iy >= g.bh || ix < 0 || ix >= g.bw) return 0.f;
        return ldf(big + ((static_cast<long long>(b) * g.bh + iy) * g.bw + ix) * g.bc + n);
    }
    __device__ void store(int m, int n, float v, int z) const {
        const int tap = z / splits;
        if (n < g.bcv) atomicAdd(dw + (static_cast<long long>(m) * g.bcv + n) * (g.k * g.k) + tap, v);
    }
};

template <class Op>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const Op op) {
    pdl_enter();
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int M = op.M(), N = op.N();
    int k_begin, k_end;
    op.krange(z, &k_begin, &k_end);

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int kt = k_begin; kt < k_end; kt += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int am, ak, bn, bk;
            if (Op::kAContigK) { ak = tid & 15; am = (tid >> 4) + 16 * i; }
            else               { am = tid & 63; ak = (tid >> 6) + 4 * i; }
            if (Op::kBContigK) { bk = tid & 15; bn = (tid >> 4) + 16 * i; }
            else               { bn = tid & 63; bk = (tid >> 6) + 4 * i; }
            As[ak][am] = (m0 + am < M && kt + ak < k_end) ? op.a(m0 + am, kt + ak, z) : 0.f;
            Bs[bk][bn] = (n0 + bn < N && kt + bk < k_end) ? op.b(kt + bk, n0 + bn, z) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) op.store(m, n, acc[i][j], z);
        }
}

Geo to_geo(const VgConvGeom* g) {
    return Geo{g->batch, g->big_h, g->big_w, g->big_c, g->small_h, g->small_w, g->small_c, g->kernel, g->stride, g->pad,
               g->big_c_valid > 0 ? g->big_c_valid : g->big_c};
}

int cdiv(long long a, int b) { return static_cast<int>((a + b - 1) / b); }

template <typename T, typename WT>
int run_down(const VgConvGeom* g, const void* big, const void* w, const float* bias, void* small, int out_f32,
             bool packed, cudaStream_t st) {
    DownOp<T, WT> op;
    op.g = to_geo(g);
    op.big = static_cast<const T*>(big);
    op.w = static_cast<const WT*>(w);
    op.bias = bias;
    op.out = small;
    op.out_f32 = out_f32;
    const long long kk = g->kernel * g->kernel;
    if (packed) { op.ws_tap = static_cast<long long>(g->small_c) * g->big_c; op.ws_sc = g->big_c; op.ws_bc = 1; }
    else        { op.ws_sc = op.g.bcv * kk; op.ws_bc = kk; op.ws_tap = 1; }
    const long long M = static_cast<long long>(g->batch) * g->small_h * g->small_w;
    dim3 grid(cdiv(M, TM), cdiv(g->small_c, TN), 1);
    launch_k(simt_gemm_kernel<decltype(op)>, dim3(grid), dim3(256), 0, st, op);
    VG_LAUNCHED();
    return VG_OK;
}

template <typename T, typename WT>
int run_up(const VgConvGeom* g, const void* small, const void* w, void* big, bool packed, cudaStream_t st) {
    UpOp<T, WT> op;
    op.g = to_geo(g);
    op.small = static_cast<const T*>(small);
    op.w = static_cast<const WT*>(w);
    op.out = static_cast<T*>(big);
    const long long kk = g->kernel * g->kernel;
    if (packed) { op.ws_tap = static_cast<long long>(g->small_c) * g->big_c; op.ws_bc = g->small_c; op.ws_sc = 1; }
    else        { op.ws_sc = op.g.bcv * kk; op.ws_bc = kk; op.ws_tap = 1; }
    const int s = g->stride;
    const long long M = static_cast<long long>(g->batch) * ((g->big_h + s - 1) / s) * ((g->big_w + s - 1) / s);
    dim3 grid(cdiv(M, TM), cdiv(g->big_c, TN), s * s);
    launch_k(simt_gemm_kernel<decltype(op)>, dim3(grid), dim3(256), 0, st, op);
    VG_LAUNCHED();
    return VG_OK;
}

template <typename T>
int run_wgrad(const VgConvGeom* g, const void* small, const void* big, float* dw, cudaStream_t st) {
    WgradOp<T> op;
    op.g = to_geo(g);
    op.small = static_cast<const T*>(small);
    op.big = static_cast<const T*>(big);
    op.dw = dw;
    const int kk = g->kernel * g->kernel;
    const long long P = static_cast<long long>(g->batch) * g->small_h * g->small_w;
    const int base = cdiv(g->small_c, TM) * cdiv(g->big_c, TN) * kk;
    int splits = std::max(1, std::min<int>(cdiv(P, 4 * TK), cdiv(148 * 4, base)));
    if (kk * splits > 65535) splits = std::max(1, 65535 / kk);
    op.splits = splits;
    dim3 grid(cdiv(g->small_c, TM), cdiv(g->big_c, TN), kk * splits);
    launch_k(simt_gemm_kernel<decltype(op)>, dim3(grid), dim3(256), 0, st, op);
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace

int simt_conv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
                   int out_f32, cudaStream_t st) {
    if (dtype == VG_F32) return run_down<float, float>(g, big, w, bias, small, out_f32, false, st);
    if (dtype == VG_BF16) return run_down<__nv_bfloat16, __nv_bfloat16>(g, big, w, bias, small, out_f32, true, st);
    return fail(VG_ERR_ARG, "down: bad dtype %d", static_cast<int>(dtype));
}

int simt_conv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, cudaStream_t st) {
    if (dtype == VG_F32) return run_up<float, float>(g, small, w, big, false, st);
    if (dtype == VG_BF16) return run_up<__nv_bfloat16, __nv_bfloat16>(g, small, w, big, true, st);
    return fail(VG_ERR_ARG, "up: bad dtype %d", static_cast<int>(dtype));
}

int simt_conv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw,
                    cudaStream_t st) {
    if (dtype == VG_F32) return run_wgrad<float>(g, small, big, dw, st);
    if (dtype == VG_BF16) return run_wgrad<__nv_bfloat16>(g, small, big, dw, st);
    return fail(VG_ERR_ARG, "wgrad: bad dtype %d", static_cast<int>(dtype));
}

}  // namespace vg
