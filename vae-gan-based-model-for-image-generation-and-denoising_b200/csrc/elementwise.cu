// HBM-bound kernels around the convolutions: BatchNorm (train statistics, apply, backward), activations,
// NCHW<->NHWC edge conversions, per-channel column sums.  All vectorised 16 B per thread access, fp32 math,
// two-stage deterministic reductions (per-block partials -> fp64 finalize), no atomics.
#include <cuda_bf16.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "pdl.cuh"
#include "vec.cuh"

namespace vg {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 4;

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
    switch (act) {
        case VG_ACT_RELU: return z > 0.f ? z : 0.f;
        case VG_ACT_LEAKY: return z > 0.f ? z : z * slope;
        case VG_ACT_TANH: return tanhf(z);
        case VG_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
        default: return z;
    }
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
    switch (act) {
        case VG_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case VG_ACT_LEAKY: return z > 0.f ? 1.f : slope;
        case VG_ACT_TANH: { const float t = tanhf(z); return 1.f - t * t; }
        case VG_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-z)); return s * (1.f - s); }
        default: return 1.f;
    }
}

// ------------------------------------------------------------------------------------------------
// Per-channel reductions over the rows of an NHWC [rows][C] tensor, finalized in the SAME launch.
//   MODE 0: (sum d, sum d^2), d = x - pivot            -> BatchNorm training statistics (+ running stats, affine)
//   MODE 1: (sum dz, sum dz*xhat), dz = dy*act'(x*scale+shift), xhat = (x-mean)*rstd   -> BatchNorm backward
//   MODE 2: (sum x)                                    -> bias gradient
// Every block reduces its row range in registers / shared memory and adds its per-channel partials (fp64 atomics) to
// a zero-initialised accumulator slot; the block that arrives last (atomic ticket) turns the totals into the
// per-channel outputs and re-zeroes the slot, so no separate finalize kernel and no memset is needed.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxChannels = 4096;
constexpr int kSlots = 64;
constexpr int kAccDoubles = 2 * kMaxChannels;
// zero at module load; every launch leaves its slot zero again.  A slot holds `reps` replicas of the [2][C] totals
// (reps = min(16, kAccDoubles / 2C)); block b adds into replica b % reps, which spreads the same-address atomics.
__device__ double g_acc[kSlots][kAccDoubles];
__device__ unsigned int g_ticket[kSlots];

struct ReduceArgs {
    const void* x;
    const void* dy;
    const float *scale, *shift, *mean, *rstd;
    long long rows;
    int C;
    long long rows_per_block;
    int act;
    float slope;
    int slot;
    // finalize, MODE 0
    const float *gamma, *beta;
    float *running_mean, *running_var;
    long long* num_batches_tracked;
    float momentum, eps;
    float *mean_out, *rstd_out, *scale_out, *shift_out;
    // finalize, MODE 1
    float *dgamma, *dbeta, *c1, *c2;
    // finalize, MODE 2
    float* colsum_out;
};

template <typename T>
__device__ __forceinline__ float load1(const T* p) {
    if constexpr (sizeof(T) == 4) return *p; else return __bfloat162float(*p);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 4) channel_reduce_kernel(const ReduceArgs a) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    __shared__ float red[kThreads][2 * V + 1];
    __shared__ bool is_last;
    const int tpr = a.C / V;                       // vectors per row (power of two)
    const int lanes = tpr < kThreads ? tpr : kThreads;
    const int rows_per_iter = kThreads / lanes;
    const int lane = threadIdx.x % lanes, rsub = threadIdx.x / lanes;
    const long long r0 = blockIdx.x * a.rows_per_block;
    const long long r1 = min(a.rows, r0 + a.rows_per_block);
    const T* x = static_cast<const T*>(a.x);
    const T* dy = static_cast<const T*>(a.dy);
    const int reps = max(1, min(16, kAccDoubles / (2 * a.C)));
    // cross-block totals: the bf16 mode adds fp32 partials with fp32 atomics on a float view of the slot; the fp32
    // (reference-accurate) mode keeps fp64 atomics
    using AccT = typename std::conditional<sizeof(T) == 4, double, float>::type;
    AccT* accf_all = reinterpret_cast<AccT*>(g_acc[a.slot]);
    AccT* acc = accf_all + static_cast<long long>(blockIdx.x % reps) * 2 * a.C;

    for (int cg = 0; cg < tpr / lanes; ++cg) {
        const int c0 = (cg * lanes + lane) * V;
        float s0[V], s1[V];
#pragma unroll
        for (int i = 0; i < V; ++i) s0[i] = s1[i] = 0.f;
        float sc[V], sh[V], mu[V], rs[V];
        float piv[V];
        if (MODE == 0) {
            // shifted-data sums: accumulate (x - K) and (x - K)^2 with K = the channel's first value, so that
            // var = E[d^2] - E[d]^2 does not cancel catastrophically when |mean| >> std or rows are few
            Vec<T>::load(x + c0, piv);
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                sc[i] = a.scale ? a.scale[c0 + i] : 1.f;
                sh[i] = a.shift ? a.shift[c0 + i] : 0.f;
                mu[i] = a.mean[c0 + i];
                rs[i] = a.rstd[c0 + i];
            }
        }
        constexpr int U = 1;   // rows in flight per thread: more costs registers -> occupancy -> waves (measured)
        for (long long r = r0 + rsub; r < r1; r += static_cast<long long>(rows_per_iter) * U) {
            float xv[U][V], dv[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long rr = r + static_cast<long long>(u) * rows_per_iter;
                if (rr < r1) {
                    Vec<T>::load(x + rr * a.C + c0, xv[u]);
                    if (MODE == 1) Vec<T>::load(dy + rr * a.C + c0, dv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long rr = r + static_cast<long long>(u) * rows_per_iter;
                if (rr >= r1) continue;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    if (MODE == 0) {
                        const float d = xv[u][i] - piv[i];
                        s0[i] += d;
                        s1[i] = fmaf(d, d, s1[i]);
                    } else if (MODE == 1) {
                        const float z = fmaf(xv[u][i], sc[i], sh[i]);
                        const float dz = dv[u][i] * act_grad(z, a.act, a.slope);
                        s0[i] += dz;
                        s1[i] = fmaf(dz, (xv[u][i] - mu[i]) * rs[i], s1[i]);
                    } else {
                        s0[i] += xv[u][i];
                    }
                }
            }
        }
        // combine the row sub-groups: inside a warp by shuffles (when a row is narrower than a warp), then across the
        // (at most 8) warps / row groups through shared memory
        if (lanes < 32) {
            for (int o = lanes; o < 32; o <<= 1) {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], o);
                    s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
                }
            }
        }
        const int groups = lanes < 32 ? kThreads / 32 : rows_per_iter;     // partial sums left per channel vector
        const int gidx = lanes < 32 ? (threadIdx.x >> 5) : rsub;
        const bool writer = lanes < 32 ? ((threadIdx.x & 31) < lanes) : true;
        if (writer) {
#pragma unroll
            for (int i = 0; i < V; ++i) { red[gidx * lanes + lane][i] = s0[i]; red[gidx * lanes + lane][V + i] = s1[i]; }
        }
        __syncthreads();
        if (threadIdx.x < lanes) {
#pragma unroll
            for (int i = 0; i < V; ++i) { s0[i] = red[lane][i]; s1[i] = red[lane][V + i]; }
            for (int j = 1; j < groups; ++j)
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    s0[i] += red[j * lanes + lane][i];
                    s1[i] += red[j * lanes + lane][V + i];
                }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                atomicAdd(acc + c0 + i, static_cast<AccT>(s0[i]));
                if (MODE != 2) atomicAdd(acc + a.C + c0 + i, static_cast<AccT>(s1[i]));
            }
        }
        __syncthreads();
    }

    // ---- ticket: the last block to arrive finalizes
    __threadfence();
    if (threadIdx.x == 0) is_last = atomicAdd(&g_ticket[a.slot], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double n = static_cast<double>(a.rows);
    AccT* all = accf_all;
    for (int c = threadIdx.x; c < a.C; c += kThreads) {
        // all replica loads are issued before the first store: interleaving them with the re-zeroing stores would
        // serialise 2*reps dependent L2 round trips (~20 us), which used to be the floor of every reduction launch
        AccT v0[16], v1[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            v0[r] = v1[r] = 0;
            if (r < reps) {
                const AccT* base = all + static_cast<long long>(r) * 2 * a.C;
                v0[r] = __ldcg(base + c);
                if (MODE != 2) v1[r] = __ldcg(base + a.C + c);
            }
        }
        double s = 0.0, ss = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) { s += static_cast<double>(v0[r]); ss += static_cast<double>(v1[r]); }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (r < reps) {
                AccT* base = all + static_cast<long long>(r) * 2 * a.C;
                base[c] = 0;
                if (MODE != 2) base[a.C + c] = 0;
            }
        }
        if (MODE == 0) {
            const double dmean = s / n;
            const double mean = static_cast<double>(load1(x + c)) + dmean;
            double var = ss / n - dmean * dmean;
            if (var < 0.0) var = 0.0;
            const double rstd = 1.0 / sqrt(var + static_cast<double>(a.eps));
            const float g = a.gamma ? a.gamma[c] : 1.f, bt = a.beta ? a.beta[c] : 0.f;
            a.mean_out[c] = static_cast<float>(mean);
            a.rstd_out[c] = static_cast<float>(rstd);
            a.scale_out[c] = static_cast<float>(g * rstd);
            a.shift_out[c] = static_cast<float>(bt - mean * g * rstd);
            if (a.running_mean != nullptr) {
                const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
                a.running_mean[c] = static_cast<float>((1.0 - a.momentum) * a.running_mean[c] + a.momentum * mean);
                a.running_var[c] = static_cast<float>((1.0 - a.momentum) * a.running_var[c] + a.momentum * unbiased);
            }
        } else if (MODE == 1) {
            if (a.dbeta != nullptr) a.dbeta[c] += static_cast<float>(s);
            if (a.dgamma != nullptr) a.dgamma[c] += static_cast<float>(ss);
            a.c1[c] = static_cast<float>(s / n);
            a.c2[c] = static_cast<float>(ss / n);
        } else {
            a.colsum_out[c] += static_cast<float>(s);
        }
    }
    if (threadIdx.x == 0) {
        g_ticket[a.slot] = 0;
        if (MODE == 0 && a.num_batches_tracked != nullptr) *a.num_batches_tracked += 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused BatchNorm passes (cooperative launch: all blocks co-resident, one grid barrier).
//   FWD : pass 1 = pivoted sums -> barrier -> every thread derives scale/shift of ITS channels from the totals ->
//         pass 2 = y = act(x*scale+shift).  Block 0 also writes mean/rstd/scale/shift and updates the running stats.
//   BWD : pass 1 = sum dz, sum dz*xhat -> barrier -> pass 2 = dx = scale*(dz - c1 - xhat*c2); block 0 adds dgamma/dbeta.
// The second pass re-reads the block's own rows, which are L2-resident for all but the largest tensors, so HBM sees
// the activation once per direction instead of twice, and two launches (reduce + apply) become one.
// Threads keep a FIXED channel group for the whole kernel, so per-channel parameters live in registers.
// ------------------------------------------------------------------------------------------------
__device__ unsigned int g_barrier[kSlots];
__device__ unsigned int g_ticket2[kSlots];

struct FusedArgs {
    const void* x;      // raw conv output
    const void* dy;     // BWD: gradient of the activated output
    void* out;          // FWD: activated output; BWD: gradient of the raw conv output
    long long rows;
    int C;
    long long rows_per_block;
    int act;
    float slope;
    int slot;
    const float *gamma, *beta;
    float *running_mean, *running_var;
    long long* num_batches_tracked;
    float momentum, eps;
    float* stats;       // [4][C]: mean, rstd, scale, shift (FWD: written, BWD: read)
    float *dgamma, *dbeta;
};

template <typename T, bool BWD>
__global__ void __launch_bounds__(kThreads) bn_fused_kernel(const FusedArgs a) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;                            // rows in flight per thread
    __shared__ float red[kThreads][2 * V + 1];
    const int lanes = a.C / V;                      // <= kThreads (checked by the launcher), power of two
    const int rows_per_iter = kThreads / lanes;
    const int lane = threadIdx.x % lanes, rsub = threadIdx.x / lanes;
    const long long r0 = blockIdx.x * a.rows_per_block;
    const long long r1 = min(a.rows, r0 + a.rows_per_block);
    const T* x = static_cast<const T*>(a.x);
    const T* dy = static_cast<const T*>(a.dy);
    T* out = static_cast<T*>(a.out);
    const int reps = max(1, min(16, kAccDoubles / (2 * a.C)));
    double* acc_all = g_acc[a.slot];
    double* acc = acc_all + static_cast<long long>(blockIdx.x % reps) * 2 * a.C;
    const int c0 = lane * V;
    const double n = static_cast<double>(a.rows);

    float sc[V], sh[V], mu[V], rs[V], piv[V];
    if (BWD) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            mu[i] = a.stats[c0 + i];
            rs[i] = a.stats[a.C + c0 + i];
            sc[i] = a.stats[2 * a.C + c0 + i];
            sh[i] = a.stats[3 * a.C + c0 + i];
        }
    } else {
        Vec<T>::load(x + c0, piv);
    }

    // ---- pass 1
    float s0[V], s1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s0[i] = s1[i] = 0.f;
    for (long long r = r0 + rsub; r < r1; r += static_cast<long long>(rows_per_iter) * U) {
        float xv[U][V], dv[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + static_cast<long long>(u) * rows_per_iter;
            if (rr < r1) {
                Vec<T>::load(x + rr * a.C + c0, xv[u]);
                if (BWD) Vec<T>::load(dy + rr * a.C + c0, dv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + static_cast<long long>(u) * rows_per_iter;
            if (rr < r1) {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    if (BWD) {
                        const float z = fmaf(xv[u][i], sc[i], sh[i]);
                        const float dz = dv[u][i] * act_grad(z, a.act, a.slope);
                        s0[i] += dz;
                        s1[i] = fmaf(dz, (xv[u][i] - mu[i]) * rs[i], s1[i]);
                    } else {
                        const float d = xv[u][i] - piv[i];
                        s0[i] += d;
                        s1[i] = fmaf(d, d, s1[i]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { red[threadIdx.x][i] = s0[i]; red[threadIdx.x][V + i] = s1[i]; }
    __syncthreads();
    if (rsub == 0) {
        for (int j = 1; j < rows_per_iter; ++j)
#pragma unroll
            for (int i = 0; i < V; ++i) {
                s0[i] += red[j * lanes + lane][i];
                s1[i] += red[j * lanes + lane][V + i];
            }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            atomicAdd(acc + c0 + i, static_cast<double>(s0[i]));
            atomicAdd(acc + a.C + c0 + i, static_cast<double>(s1[i]));
        }
    }

    // ---- grid barrier (co-residency is guaranteed by the cooperative launch)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&g_barrier[a.slot], 1u);
        const long long t_start = clock64();
        while (*reinterpret_cast<volatile unsigned int*>(&g_barrier[a.slot]) < gridDim.x) {
            if (clock64() - t_start > 4000000000LL) {   // ~2 s: turn a would-be hang into a CUDA error
                printf("vg: bn_fused_kernel grid barrier timed out (block %d of %d)\n", blockIdx.x, gridDim.x);
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();

    // ---- totals of this thread's channels
    float c1[V], c2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        double t0 = 0.0, t1 = 0.0;
        for (int r = 0; r < reps; ++r) {
            t0 += __ldcg(acc_all + static_cast<long long>(r) * 2 * a.C + c0 + i);
            t1 += __ldcg(acc_all + static_cast<long long>(r) * 2 * a.C + a.C + c0 + i);
        }
        if (BWD) {
            c1[i] = static_cast<float>(t0 / n);
            c2[i] = static_cast<float>(t1 / n);
            if (blockIdx.x == 0 && rsub == 0) {
                if (a.dbeta != nullptr) a.dbeta[c0 + i] += static_cast<float>(t0);
                if (a.dgamma != nullptr) a.dgamma[c0 + i] += static_cast<float>(t1);
            }
        } else {
            const double dmean = t0 / n;
            const double mean = static_cast<double>(piv[i]) + dmean;
            double var = t1 / n - dmean * dmean;
            if (var < 0.0) var = 0.0;
            const double rstd = 1.0 / sqrt(var + static_cast<double>(a.eps));
            const float g = a.gamma ? a.gamma[c0 + i] : 1.f, bt = a.beta ? a.beta[c0 + i] : 0.f;
            sc[i] = static_cast<float>(g * rstd);
            sh[i] = static_cast<float>(bt - mean * g * rstd);
            if (blockIdx.x == 0 && rsub == 0) {
                const int c = c0 + i;
                a.stats[c] = static_cast<float>(mean);
                a.stats[a.C + c] = static_cast<float>(rstd);
                a.stats[2 * a.C + c] = sc[i];
                a.stats[3 * a.C + c] = sh[i];
                if (a.running_mean != nullptr) {
                    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
                    a.running_mean[c] = static_cast<float>((1.0 - a.momentum) * a.running_mean[c] + a.momentum * mean);
                    a.running_var[c] = static_cast<float>((1.0 - a.momentum) * a.running_var[c] + a.momentum * unbiased);
                }
            }
        }
    }
    if (!BWD && blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches_tracked != nullptr) *a.num_batches_tracked += 1;

    // ---- everyone has read the totals once the second ticket is full: the last block re-arms the slot
    __shared__ bool last2;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last2 = atomicAdd(&g_ticket2[a.slot], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last2) {
        for (int i = threadIdx.x; i < reps * 2 * a.C; i += kThreads) acc_all[i] = 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            g_barrier[a.slot] = 0;
            g_ticket2[a.slot] = 0;
            __threadfence();
        }
    }

    // ---- pass 2
    for (long long r = r0 + rsub; r < r1; r += static_cast<long long>(rows_per_iter) * U) {
        float xv[U][V], dv[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + static_cast<long long>(u) * rows_per_iter;
            if (rr < r1) {
                Vec<T>::load(x + rr * a.C + c0, xv[u]);
                if (BWD) Vec<T>::load(dy + rr * a.C + c0, dv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + static_cast<long long>(u) * rows_per_iter;
            if (rr < r1) {
                float o[V];
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float z = fmaf(xv[u][i], sc[i], sh[i]);
                    if (BWD) {
                        const float dz = dv[u][i] * act_grad(z, a.act, a.slope);
                        const float xhat = (xv[u][i] - mu[i]) * rs[i];
                        o[i] = sc[i] * (dz - c1[i] - xhat * c2[i]);
                    } else {
                        o[i] = act_fwd(z, a.act, a.slope);
                    }
                }
                Vec<T>::store(out + rr * a.C + c0, o);
            }
        }
    }
}

struct ReducePlan {
    int blocks;
    long long rows_per_block;
};
// `rows_per_iter` = rows one block covers per loop iteration (256 threads / vectors per row): small tensors get many
// short blocks (2 rows per thread) so that the pass is not bound by a few threads' dependent load chains.
ReducePlan plan_reduce(long long rows, int rows_per_iter = 16) {
    ReducePlan p;
    const long long per_block = static_cast<long long>(std::max(1, rows_per_iter)) * 2;
    p.blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(kMaxBlocks, (rows + per_block - 1) / per_block)));
    p.rows_per_block = (rows + p.blocks - 1) / p.blocks;
    p.blocks = static_cast<int>((rows + p.rows_per_block - 1) / p.rows_per_block);
    return p;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int check_channels(VgDType dt, int C) {
    const int V = dt == VG_BF16 ? 8 : 4;
    if (C % V != 0 || !is_pow2(C / V) || C > kMaxChannels)
        return fail(VG_ERR_SHAPE, "channel count %d must be %d x a power of two (<= %d) for the per-channel kernels", C,
                    V, kMaxChannels);
    return VG_OK;
}

int next_slot() {
    static std::atomic<unsigned> counter{0};
    return static_cast<int>(counter.fetch_add(1, std::memory_order_relaxed) % kSlots);
}

template <int MODE>
int launch_reduce(VgDType dt, ReduceArgs a, int blocks, cudaStream_t st) {
    a.slot = next_slot();
    if (dt == VG_BF16) launch_k(channel_reduce_kernel<__nv_bfloat16, MODE>, dim3(blocks), dim3(kThreads), 0, st, a);
    else launch_k(channel_reduce_kernel<float, MODE>, dim3(blocks), dim3(kThreads), 0, st, a);
    VG_LAUNCHED();
    return VG_OK;
}

// Cooperative launch of the fused BatchNorm kernel; returns VG_ERR_SHAPE when the shape does not fit (caller falls
// back to the two-kernel path).
template <bool BWD>
int launch_fused(VgDType dt, FusedArgs a, cudaStream_t st) {
    const int V = dt == VG_BF16 ? 8 : 4;
    if (a.C % V != 0 || !is_pow2(a.C / V) || a.C / V > kThreads || a.C > kMaxChannels) return VG_ERR_SHAPE;
    static int occ[2][2] = {{0, 0}, {0, 0}};
    int& o = occ[dt == VG_BF16][BWD];
    if (o == 0) {
        int v = 0;
        cudaError_t e = dt == VG_BF16
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, bn_fused_kernel<__nv_bfloat16, BWD>, kThreads, 0)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, bn_fused_kernel<float, BWD>, kThreads, 0);
        if (e != cudaSuccess || v <= 0) return VG_ERR_SHAPE;
        o = v;
    }
    int sms = 148;
    const int lanes = a.C / V, rows_per_iter = kThreads / lanes;
    long long want = (a.rows + static_cast<long long>(rows_per_iter) * 4 - 1) / (static_cast<long long>(rows_per_iter) * 4);
    int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(want, static_cast<long long>(sms) * std::min(o, 4))));
    a.rows_per_block = (a.rows + blocks - 1) / blocks;
    blocks = static_cast<int>((a.rows + a.rows_per_block - 1) / a.rows_per_block);
    a.slot = next_slot();
    void* args[] = {&a};
    cudaError_t e = dt == VG_BF16
        ? cudaLaunchCooperativeKernel(reinterpret_cast<void*>(bn_fused_kernel<__nv_bfloat16, BWD>), dim3(blocks), dim3(kThreads), args, 0, st)
        : cudaLaunchCooperativeKernel(reinterpret_cast<void*>(bn_fused_kernel<float, BWD>), dim3(blocks), dim3(kThreads), args, 0, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel(bn_fused_kernel)");
    note_launch();
    return VG_OK;
}

__global__ void bn_eval_coeffs_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                      int C, float* scale, float* shift) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float rstd = 1.f / sqrtf(rv[c] + eps);
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    scale[c] = g * rstd;
    shift[c] = bt - rm[c] * g * rstd;
}

template <typename T>
__global__ void colsum_generic_kernel(const T* __restrict__ x, long long rows, int C, float* __restrict__ out) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (long long r = 0; r < rows; ++r) {
        if constexpr (sizeof(T) == 4) s += x[r * C + c]; else s += __bfloat162float(x[r * C + c]);
    }
    out[c] += static_cast<float>(s);
}

// ---- elementwise passes
template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads) scale_shift_act_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                                  long long n, int C, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, int act, float slope) {
    pdl_enter();
    // generic scalar path (used for tiny tensors such as the [B] discriminator logits)
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        float v;
        if constexpr (sizeof(TI) == 4) v = x[i]; else v = __bfloat162float(x[i]);
        v = fmaf(v, scale ? scale[c] : 1.f, shift ? shift[c] : 0.f);
        v = act_fwd(v, act, slope);
        if constexpr (sizeof(TO) == 4) y[i] = v; else y[i] = __float2bfloat16_rn(v);
    }
}

// The two apply passes are launched with a thread count that is a multiple of the vectors per row, so every
// thread keeps ONE channel group for its whole grid-stride loop: per-channel parameters are loaded once into
// registers, and 4 independent 16-byte loads are in flight per tensor.
template <typename T>
__global__ void __launch_bounds__(kThreads) scale_shift_act_vec_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                      long long nvec, int C,
                                                                      const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, int act,
                                                                      float slope) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int c0 = static_cast<int>((tid * V) % C);
    float sc[V], sh[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        sc[j] = scale ? __ldg(scale + c0 + j) : 1.f;
        sh[j] = shift ? __ldg(shift + c0 + j) : 0.f;
    }
    for (long long i = tid; i < nvec; i += stride * U) {
        float v[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < nvec) Vec<T>::load(x + (i + u * stride) * V, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) continue;
#pragma unroll
            for (int j = 0; j < V; ++j) v[u][j] = act_fwd(fmaf(v[u][j], sc[j], sh[j]), act, slope);
            Vec<T>::store(y + (i + u * stride) * V, v[u]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bn_act_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                   T* __restrict__ dx, long long nvec, int C,
                                                                   const float* __restrict__ scale,
                                                                   const float* __restrict__ shift,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd,
                                                                   const float* __restrict__ c1,
                                                                   const float* __restrict__ c2, int act, float slope) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int c0 = static_cast<int>((tid * V) % C);
    // dx = scale*(dz - c1 - xhat*c2) = scale*dz + kx*x + k0   with xhat = (x - mean)*rstd
    float sc[V], sh[V], kx[V], k0[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = c0 + j;
        sc[j] = __ldg(scale + c);
        sh[j] = __ldg(shift + c);
        const float t = sc[j] * __ldg(c2 + c) * __ldg(rstd + c);
        kx[j] = -t;
        k0[j] = t * __ldg(mean + c) - sc[j] * __ldg(c1 + c);
    }
    for (long long i = tid; i < nvec; i += stride * U) {
        float xv[U][V], dv[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < nvec) {
                Vec<T>::load(x + (i + u * stride) * V, xv[u]);
                Vec<T>::load(dy + (i + u * stride) * V, dv[u]);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) continue;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float z = fmaf(xv[u][j], sc[j], sh[j]);
                const float dz = dv[u][j] * act_grad(z, act, slope);
                dv[u][j] = fmaf(sc[j], dz, fmaf(kx[j], xv[u][j], k0[j]));
            }
            Vec<T>::store(dx + (i + u * stride) * V, dv[u]);
        }
    }
}

// ---- BatchNorm passes fed by the raw per-channel sums a convolution epilogue accumulated (VgEpilogue modes 1 / 2).
// The finalisation (sums -> mean, rstd, scale, shift, or -> the two projection coefficients) is a few flops per
// channel, so every thread redoes it for ITS channel group in the prologue instead of paying for a finalize launch;
// block (0, group) publishes the per-channel results, block (0, 0) walks the running statistics through all groups
// in order (momentum updates do not commute).  blockIdx.y = statistics group (sub-batch of `rows` rows).
template <typename T>
__global__ void __launch_bounds__(kThreads, 3) bn_apply_from_sums_kernel(
    const T* __restrict__ x, T* __restrict__ y, long long nvec, int C, long long rows, const float* __restrict__ sums,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
    long long* num_batches_tracked, float momentum, float eps, float* __restrict__ stats, int act, float slope) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    const int grp = blockIdx.y, groups = gridDim.y;
    x += static_cast<long long>(grp) * nvec * V;
    y += static_cast<long long>(grp) * nvec * V;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int c0 = static_cast<int>((tid * V) % C);
    const double inv_n = 1.0 / static_cast<double>(rows);
    const bool publisher = blockIdx.x == 0 && threadIdx.x < C / V;
    // the first batch of activation loads is issued BEFORE the finalisation below waits on the sums: short passes were
    // paying a full dependent round trip (sums -> math -> first load) on top of ~3 us of streaming
    typename Vec<T>::Raw raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (tid + u * stride < nvec) raw[u] = Vec<T>::load_raw(x + (tid + u * stride) * V);
    float sc[V], sh[V];
    const float* sg = sums + static_cast<long long>(grp) * 2 * C;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = c0 + j;
        // The fp32 sums carry ~1e-7 relative error already; a double-precision divide + square root per channel in every
        // thread's prologue was a multi-microsecond latency in front of each (short) apply pass.  fp32 with one
        // Newton step on rsqrt is within 2 ulp of the double result.
        // E[x^2] - mean^2 is formed in double (three fp64 operations per channel): in fp32 the subtraction itself
        // loses log2(mean^2 / var) bits on layers whose |mean| >> std (biased encoder blocks); the fp32 sums are the
        // remaining error source
        const double m = static_cast<double>(sg[c]) * inv_n;
        const float mf = static_cast<float>(m);
        const float var_f = static_cast<float>(fmax(fma(-m, m, static_cast<double>(sg[C + c]) * inv_n), 0.0));
        float rstd = rsqrtf(var_f + eps);
        rstd = rstd * fmaf(-0.5f * (var_f + eps) * rstd, rstd, 1.5f);
        const float g = gamma ? __ldg(gamma + c) : 1.f, bt = beta ? __ldg(beta + c) : 0.f;
        sc[j] = g * rstd;
        sh[j] = bt - mf * sc[j];
    }
    if (publisher) {
        // (block (0, group) only, kept out of the unrolled loop above: its fp64 running-statistics walk would otherwise
        // set the register count of every thread of the pass)
        float* st = stats + static_cast<long long>(grp) * 4 * C;
#pragma unroll 1
        for (int j = 0; j < V; ++j) {
            const int c = c0 + j;
            const double m = static_cast<double>(sg[c]) * inv_n;
            const float var_f = static_cast<float>(fmax(fma(-m, m, static_cast<double>(sg[C + c]) * inv_n), 0.0));
            float rstd = rsqrtf(var_f + eps);
            rstd = rstd * fmaf(-0.5f * (var_f + eps) * rstd, rstd, 1.5f);
            const float g = gamma ? __ldg(gamma + c) : 1.f, bt = beta ? __ldg(beta + c) : 0.f;
            const float scj = g * rstd;
            st[c] = static_cast<float>(m);
            st[C + c] = rstd;
            st[2 * C + c] = scj;
            st[3 * C + c] = bt - static_cast<float>(m) * scj;
            if (grp == 0 && running_mean != nullptr) {
                const double n = static_cast<double>(rows);
                float rm = running_mean[c], rv = running_var[c];
                for (int q = 0; q < groups; ++q) {
                    const float* sq = sums + static_cast<long long>(q) * 2 * C;
                    const double mq = static_cast<double>(sq[c]) * inv_n;
                    double vq = static_cast<double>(sq[C + c]) * inv_n - mq * mq;
                    if (vq < 0.0) vq = 0.0;
                    const double unbiased = n > 1.0 ? vq * n / (n - 1.0) : vq;
                    rm = static_cast<float>((1.0 - momentum) * rm + momentum * mq);
                    rv = static_cast<float>((1.0 - momentum) * rv + momentum * unbiased);
                }
                running_mean[c] = rm;
                running_var[c] = rv;
            }
        }
    }
    if (blockIdx.x == 0 && grp == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += groups;
    for (long long i = tid; i < nvec; i += stride * U) {
        if (i != tid) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride < nvec) raw[u] = Vec<T>::load_raw(x + (i + u * stride) * V);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) continue;
            float v[V];
            Vec<T>::unpack(raw[u], v);
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = act_fwd(fmaf(v[j], sc[j], sh[j]), act, slope);
            Vec<T>::store(y + (i + u * stride) * V, v);
        }
    }
}

// dx = scale*(dz - c1 - xhat*c2) with c1 = sum(dz)/n, c2 = sum(dz*xhat)/n taken from `sums`; dz already carries the
// activation derivative (VG_EPI_BN_BWD).  dgamma += sum(dz*xhat), dbeta += sum(dz) (atomics: groups share them).
template <typename T>
__global__ void __launch_bounds__(kThreads, 3) bn_bwd_apply_from_sums_kernel(
    const T* __restrict__ dz, const T* __restrict__ x, T* __restrict__ dx, long long nvec, int C, long long rows,
    const float* __restrict__ stats, const float* __restrict__ sums, float* dgamma, float* dbeta) {
    pdl_enter();
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    const int grp = blockIdx.y;
    dz += static_cast<long long>(grp) * nvec * V;
    x += static_cast<long long>(grp) * nvec * V;
    dx += static_cast<long long>(grp) * nvec * V;
    stats += static_cast<long long>(grp) * 4 * C;
    sums += static_cast<long long>(grp) * 2 * C;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int c0 = static_cast<int>((tid * V) % C);
    const float inv_n = 1.f / static_cast<float>(rows);
    const bool publisher = blockIdx.x == 0 && threadIdx.x < C / V;
    typename Vec<T>::Raw xr[U], dr[U];   // first batch in flight (packed) while the per-channel coefficients are fetched
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (tid + u * stride < nvec) {
            xr[u] = Vec<T>::load_raw(x + (tid + u * stride) * V);
            dr[u] = Vec<T>::load_raw(dz + (tid + u * stride) * V);
        }
    float sc[V], kx[V], k0[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = c0 + j;
        const float s0 = sums[c], s1 = sums[C + c];
        sc[j] = __ldg(stats + 2 * C + c);
        const float t = sc[j] * (s1 * inv_n) * __ldg(stats + C + c);
        kx[j] = -t;
        k0[j] = t * __ldg(stats + c) - sc[j] * (s0 * inv_n);
        if (publisher) {
            if (dbeta != nullptr) atomicAdd(dbeta + c, s0);
            if (dgamma != nullptr) atomicAdd(dgamma + c, s1);
        }
    }
    for (long long i = tid; i < nvec; i += stride * U) {
        if (i != tid) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride < nvec) {
                    xr[u] = Vec<T>::load_raw(x + (i + u * stride) * V);
                    dr[u] = Vec<T>::load_raw(dz + (i + u * stride) * V);
                }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride >= nvec) continue;
            float xv[V], dv[V];
            Vec<T>::unpack(xr[u], xv);
            Vec<T>::unpack(dr[u], dv);
#pragma unroll
            for (int j = 0; j < V; ++j) dv[j] = fmaf(sc[j], dv[j], fmaf(kx[j], xv[j], k0[j]));
            Vec<T>::store(dx + (i + u * stride) * V, dv);
        }
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads) act_bwd_kernel(const TI* __restrict__ dy, const TI* __restrict__ x,
                                                          TO* __restrict__ dx, long long n, int act, float slope) {
    pdl_enter();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float d, z;
        if constexpr (sizeof(TI) == 4) { d = dy[i]; z = x[i]; }
        else { d = __bfloat162float(dy[i]); z = __bfloat162float(x[i]); }
        const float v = d * act_grad(z, act, slope);
        if constexpr (sizeof(TO) == 4) dx[i] = v; else dx[i] = __float2bfloat16_rn(v);
    }
}

// ---- NCHW fp32 <-> NHWC T at the module edges.  One thread per (b, h, w) pixel: reads/writes of the NCHW side are
// coalesced along w; the NHWC side is C contiguous elements per thread.
//   mode 0: dst = src                         mode 1: dst = clamp?(src + sigma * aux)
//   mode 2: dst = src * (1 - aux^2)           (tanh backward; aux = tanh output, NCHW)
template <typename T>
__global__ void __launch_bounds__(kThreads) nchw_to_nhwc_kernel(const float* __restrict__ src,
                                                               const float* __restrict__ aux, T* __restrict__ dst,
                                                               int B, int C, int Cd, long long HW, int mode,
                                                               float sigma, int clamp) {
    pdl_enter();
    const long long total = static_cast<long long>(B) * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        if (sizeof(T) == 2 && Cd == 16 && C <= 16) {
            // padded 16-channel bf16 pixel: 32 contiguous bytes, two 16-byte stores
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                v[c] = 0.f;
                if (c < C) {
                    const long long s = (b * C + c) * HW + p;
                    float t = src[s];
                    if (mode == 1) {
                        t = fmaf(sigma, aux[s], t);
                        if (clamp) t = fminf(1.f, fmaxf(-1.f, t));
                    } else if (mode == 2) {
                        const float y = aux[s];
                        t *= (1.f - y * y);
                    }
                    v[c] = t;
                }
            }
            float lo[8], hi[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { lo[c] = v[c]; hi[c] = v[8 + c]; }
            Vec<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(dst) + i * 16, lo);
            Vec<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(dst) + i * 16 + 8, hi);
            continue;
        }
        for (int c = 0; c < Cd; ++c) {
            float v = 0.f;
            if (c < C) {
                const long long s = (b * C + c) * HW + p;
                v = src[s];
                if (mode == 1) {
                    v = fmaf(sigma, aux[s], v);
                    if (clamp) v = fminf(1.f, fmaxf(-1.f, v));
                } else if (mode == 2) {
                    const float y = aux[s];
                    v *= (1.f - y * y);
                }
            }
            if constexpr (sizeof(T) == 4) dst[i * Cd + c] = v; else dst[i * Cd + c] = __float2bfloat16_rn(v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                                               int B, int C, int Cs, long long HW, int act,
                                                               float slope) {
    pdl_enter();
    const long long total = static_cast<long long>(B) * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        for (int c = 0; c < C; ++c) {
            float v;
            if constexpr (sizeof(T) == 4) v = src[i * Cs + c]; else v = __bfloat162float(src[i * Cs + c]);
            dst[(b * C + c) * HW + p] = act_fwd(v, act, slope);
        }
    }
}

// ---- space-to-depth ("s2d") image tensors.  A C <= 16 channel H x W image is held as bf16 [B][H/2+o][W/2+o][64]:
// block (Y, X), sub-pixel (sy, sx), channel c sits at slot (sy*2+sx)*16 + c and is pixel (2Y-o+sy, 2X-o+sx); every other
// slot is zero.  With 128-byte rows the stride-2 4x4 image convolutions become plain 2x2 stride-1 convolutions over 64
// channels (origin o = the convolution's padding), which the TMA-fed tensor-core kernels run at full row efficiency.
__global__ void __launch_bounds__(kThreads) nchw_to_s2d_kernel(const float* __restrict__ src,
                                                              const float* __restrict__ aux,
                                                              __nv_bfloat16* __restrict__ dst, int B, int C, int H, int W,
                                                              int o, int mode, float sigma, int clamp) {
    pdl_enter();
    const int bh = H / 2 + o, bw = W / 2 + o;
    const long long total = static_cast<long long>(B) * bh * bw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // (32-bit index arithmetic whenever the block count allows it: three 64-bit divisions per thread cost more
        // instructions than the rest of this kernel)
        int X, Y;
        long long b;
        if (total <= 0x7fffffffLL) {
            const unsigned ii = static_cast<unsigned>(i), q = ii / static_cast<unsigned>(bw);
            X = static_cast<int>(ii - q * bw);
            Y = static_cast<int>(q % static_cast<unsigned>(bh));
            b = q / static_cast<unsigned>(bh);
        } else {
            X = static_cast<int>(i % bw);
            Y = static_cast<int>((i / bw) % bh);
            b = i / (static_cast<long long>(bw) * bh);
        }
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
            const int py = 2 * Y - o + (sub >> 1), px = 2 * X - o + (sub & 1);
            const bool in = py >= 0 && py < H && px >= 0 && px < W;
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                v[c] = 0.f;
                if (in && c < C) {
                    const long long s = ((b * C + c) * H + py) * W + px;
                    float t = src[s];
                    if (mode == 1) {
                        t = fmaf(sigma, aux[s], t);
                        if (clamp) t = fminf(1.f, fmaxf(-1.f, t));
                    } else if (mode == 2) {
                        const float y = aux[s];
                        t *= (1.f - y * y);
                    }
                    v[c] = t;
                }
            }
            float lo[8], hi[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { lo[c] = v[c]; hi[c] = v[8 + c]; }
            Vec<__nv_bfloat16>::store(dst + i * 64 + sub * 16, lo);
            Vec<__nv_bfloat16>::store(dst + i * 64 + sub * 16 + 8, hi);
        }
    }
}

__global__ void __launch_bounds__(kThreads) s2d_to_nchw_kernel(const __nv_bfloat16* __restrict__ src,
                                                              float* __restrict__ dst, int B, int C, int H, int W, int o,
                                                              int act, float slope) {
    pdl_enter();
    const int bh = H / 2 + o, bw = W / 2 + o;
    const long long total = static_cast<long long>(B) * bh * bw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // (32-bit index arithmetic whenever the block count allows it: three 64-bit divisions per thread cost more
        // instructions than the rest of this kernel)
        int X, Y;
        long long b;
        if (total <= 0x7fffffffLL) {
            const unsigned ii = static_cast<unsigned>(i), q = ii / static_cast<unsigned>(bw);
            X = static_cast<int>(ii - q * bw);
            Y = static_cast<int>(q % static_cast<unsigned>(bh));
            b = q / static_cast<unsigned>(bh);
        } else {
            X = static_cast<int>(i % bw);
            Y = static_cast<int>((i / bw) % bh);
            b = i / (static_cast<long long>(bw) * bh);
        }
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
            const int py = 2 * Y - o + (sub >> 1), px = 2 * X - o + (sub & 1);
            if (py < 0 || py >= H || px < 0 || px >= W) continue;
            float v[8];
            Vec<__nv_bfloat16>::load(src + i * 64 + sub * 16, v);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < C) dst[((b * C + c) * H + py) * W + px] = act_fwd(v[c], act, slope);
            if (C > 8) {
                Vec<__nv_bfloat16>::load(src + i * 64 + sub * 16 + 8, v);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (8 + c < C) dst[((b * C + 8 + c) * H + py) * W + px] = act_fwd(v[c], act, slope);
            }
        }
    }
}

// uint8 NHWC image batch (what an image decoder produces; dataset_code.py:147-150 then applies ToTensor + Normalize)
// -> fp32 NCHW, y = (x/255 - mean) / std.  One thread per pixel: coalesced 3-byte reads, W-contiguous writes.
__global__ void __launch_bounds__(kThreads) u8_nhwc_to_nchw_kernel(const unsigned char* __restrict__ src,
                                                                  float* __restrict__ dst, int B, int C, long long HW,
                                                                  float mean, float inv_std) {
    pdl_enter();
    const long long total = static_cast<long long>(B) * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        for (int c = 0; c < C; ++c)
            dst[(b * C + c) * HW + p] = (static_cast<float>(src[i * C + c]) * (1.f / 255.f) - mean) * inv_std;
    }
}

// nn.Linear over a flattened NCHW feature map (main_vae.py:47-48,53) as ONE dense GEMM over the NHWC activation:
// the master W[n][c*kk + tap] (reference (c, h, w) flatten order) and the GEMM operand W'[n][tap*C + c] ((h, w, c) order,
// rows n_valid..n_pad-1 zero) differ by a per-row [C][kk] <-> [kk][C] transpose.  32x32 shared-memory tiles, both
// sides coalesced.  mode 0: dst = W' built from src = W;  mode 1: dst = dW (master layout) += src = dW'.
__global__ void __launch_bounds__(256) linear_permute_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int n_valid, int C, int kk, int mode) {
    pdl_enter();
    __shared__ float tile[32][33];
    const int n = blockIdx.z, tap0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
    const long long row = static_cast<long long>(n) * C * kk;
    if (mode == 0) {
        for (int r = ty; r < 32; r += 8) {                             // read W[n][c0+r][tap0+tx]
            const int c = c0 + r, tap = tap0 + tx;
            tile[r][tx] = (n < n_valid && c < C && tap < kk) ? src[row + static_cast<long long>(c) * kk + tap] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {                             // write W'[n][(tap0+r)*C + c0+tx]
            const int tap = tap0 + r, c = c0 + tx;
            if (tap < kk && c < C) dst[row + static_cast<long long>(tap) * C + c] = tile[tx][r];
        }
    } else {
        if (n >= n_valid) return;
        for (int r = ty; r < 32; r += 8) {                             // read dW'[n][(tap0+r)*C + c0+tx]
            const int tap = tap0 + r, c = c0 + tx;
            tile[r][tx] = (tap < kk && c < C) ? src[row + static_cast<long long>(tap) * C + c] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {                             // dW[n][(c0+r)*kk + tap0+tx] +=
            const int c = c0 + r, tap = tap0 + tx;
            if (c < C && tap < kk) dst[row + static_cast<long long>(c) * kk + tap] += tile[tx][r];
        }
    }
}

// dst[i] (+)= sum_j src[idx[i*fan + j]]  (idx < 0 = no term): equivalent-weight construction and its gradient
__global__ void __launch_bounds__(kThreads) gather_f32_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                             const int* __restrict__ idx, long long n, int fan,
                                                             int accumulate) {
    pdl_enter();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float acc = accumulate ? dst[i] : 0.f;
        for (int j = 0; j < fan; ++j) {
            const int k = __ldg(idx + i * fan + j);
            if (k >= 0) acc += __ldg(src + k);
        }
        dst[i] = acc;
    }
}

int grid_for(long long n) {
    return static_cast<int>(std::max<long long>(1, std::min<long long>(148 * 8, (n + kThreads - 1) / kThreads)));
}

// Grid for the channel-affine apply kernels: ~nvec/4 threads, total thread count a multiple of `vec_per_row`.
// block cap of the from-sums BatchNorm passes: one resident wave split over the statistics groups (experiment switch
// VG_BN_WAVE=0: the older 148 * 8 cap, several waves of short-lived blocks)
int bn_wave_blocks(int groups) {
    static const bool wave = [] {
        const char* v = getenv("VG_BN_WAVE");
        return v == nullptr || v[0] != '0';
    }();
    return wave ? std::max(1, 148 * 3 / groups) : 148 * 8;
}

int grid_affine(long long nvec, int vec_per_row, int max_blocks = 148 * 8) {
    long long blocks = std::max<long long>(1, std::min<long long>(max_blocks, (nvec / 4 + kThreads - 1) / kThreads));
    const int mult = std::max(1, vec_per_row / kThreads);          // vec_per_row is a power of two
    blocks = (blocks + mult - 1) / mult * mult;
    return static_cast<int>(blocks);
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" size_t vg_reduce_workspace_bytes(long long rows, int channels) {
    (void)rows;
    return static_cast<size_t>(channels) * sizeof(float);   // kept for ABI stability; the reductions need no scratch
}

extern "C" int vg_bn_train_fwd(const void* x, VgDType dt, long long rows, int C, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float momentum, float eps, float* mean_out,
                               float* rstd_out, float* scale_out, float* shift_out, float* ws, size_t ws_bytes,
                               void* stream) {
    (void)ws;
    (void)ws_bytes;
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (x == nullptr || mean_out == nullptr || rstd_out == nullptr || scale_out == nullptr || shift_out == nullptr)
        return fail(VG_ERR_ARG, "bn_train_fwd: null pointer");
    rc = check_channels(dt, C);
    if (rc != VG_OK) return rc;
    if (rows < 2) return fail(VG_ERR_SHAPE, "Expected more than 1 value per channel when training");
    const ReducePlan p = plan_reduce(rows, std::max(1, kThreads / (C / (dt == VG_BF16 ? 8 : 4))));
    ReduceArgs a{};
    a.x = x; a.rows = rows; a.C = C; a.rows_per_block = p.rows_per_block;
    a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var;
    a.num_batches_tracked = num_batches_tracked; a.momentum = momentum; a.eps = eps;
    a.mean_out = mean_out; a.rstd_out = rstd_out; a.scale_out = scale_out; a.shift_out = shift_out;
    return launch_reduce<0>(dt, a, p.blocks, as_stream(stream));
}

extern "C" int vg_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, float eps, int C, float* scale_out, float* shift_out,
                                 void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    launch_k(bn_eval_coeffs_kernel, dim3((C + 127) / 128), dim3(128), 0, as_stream(stream), gamma, beta, running_mean, running_var, eps,
                                                                          C, scale_out, shift_out);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_scale_shift_act(const void* x, VgDType in_dt, long long rows, int C, const float* scale,
                                  const float* shift, VgAct act, float slope, void* y, VgDType out_dt, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (x == nullptr || y == nullptr) return fail(VG_ERR_ARG, "scale_shift_act: null pointer");
    const long long n = rows * C;
    cudaStream_t st = as_stream(stream);
    const int V = in_dt == VG_BF16 ? 8 : 4;
    if (in_dt == out_dt && C % V == 0 && is_pow2(C / V)) {
        const long long nvec = n / V;
        if (in_dt == VG_BF16)
            launch_k(scale_shift_act_vec_kernel<__nv_bfloat16>, dim3(grid_affine(nvec, C / V)), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), nvec, C, scale, shift, act, slope);
        else
            launch_k(scale_shift_act_vec_kernel<float>, dim3(grid_affine(nvec, C / V)), dim3(kThreads), 0, st, static_cast<const float*>(x), static_cast<float*>(y), nvec, C, scale, shift, act, slope);
    } else if (in_dt == VG_BF16 && out_dt == VG_BF16) {
        launch_k(scale_shift_act_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid_for(n)), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, C, scale, shift, act, slope);
    } else if (in_dt == VG_BF16 && out_dt == VG_F32) {
        launch_k(scale_shift_act_kernel<__nv_bfloat16, float>, dim3(grid_for(n)), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(x), static_cast<float*>(y), n, C, scale, shift, act, slope);
    } else if (in_dt == VG_F32 && out_dt == VG_BF16) {
        launch_k(scale_shift_act_kernel<float, __nv_bfloat16>, dim3(grid_for(n)), dim3(kThreads), 0, st, static_cast<const float*>(x), static_cast<__nv_bfloat16*>(y), n, C, scale, shift, act, slope);
    } else {
        launch_k(scale_shift_act_kernel<float, float>, dim3(grid_for(n)), dim3(kThreads), 0, st, static_cast<const float*>(x), static_cast<float*>(y), n, C, scale, shift, act, slope);
    }
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_bn_act_bwd(const void* dy, const void* x, VgDType dt, long long rows, int C, const float* scale,
                             const float* shift, const float* mean, const float* rstd, VgAct act, float slope,
                             float* dgamma, float* dbeta, void* dx, float* ws, size_t ws_bytes, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (dy == nullptr || x == nullptr || dx == nullptr || scale == nullptr || shift == nullptr || mean == nullptr ||
        rstd == nullptr)
        return fail(VG_ERR_ARG, "bn_act_bwd: null pointer");
    rc = check_channels(dt, C);
    if (rc != VG_OK) return rc;
    const ReducePlan p = plan_reduce(rows, std::max(1, kThreads / (C / (dt == VG_BF16 ? 8 : 4))));
    const size_t need = 2 * static_cast<size_t>(C) * sizeof(float);
    if (ws == nullptr || ws_bytes < need) return fail(VG_ERR_WORKSPACE, "bn_act_bwd: workspace too small");
    float* c1 = ws;
    float* c2 = c1 + C;
    cudaStream_t st = as_stream(stream);
    ReduceArgs a{};
    a.x = x; a.dy = dy; a.scale = scale; a.shift = shift; a.mean = mean; a.rstd = rstd;
    a.rows = rows; a.C = C; a.rows_per_block = p.rows_per_block; a.act = act; a.slope = slope;
    a.dgamma = dgamma; a.dbeta = dbeta; a.c1 = c1; a.c2 = c2;
    rc = launch_reduce<1>(dt, a, p.blocks, st);
    if (rc != VG_OK) return rc;
    const int V = dt == VG_BF16 ? 8 : 4;
    const long long nvec = rows * C / V;
    if (dt == VG_BF16)
        launch_k(bn_act_bwd_apply_kernel<__nv_bfloat16>, dim3(grid_affine(nvec, C / V)), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
            static_cast<__nv_bfloat16*>(dx), nvec, C, scale, shift, mean, rstd, c1, c2, act, slope);
    else
        launch_k(bn_act_bwd_apply_kernel<float>, dim3(grid_affine(nvec, C / V)), dim3(kThreads), 0, st, static_cast<const float*>(dy), static_cast<const float*>(x), static_cast<float*>(dx), nvec, C, scale, shift,
            mean, rstd, c1, c2, act, slope);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_bn_apply_from_sums(const void* x, VgDType dt, long long rows, int C, int groups, const float* sums,
                                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, float momentum, float eps, VgAct act, float slope,
                                     float* stats, void* y, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (x == nullptr || y == nullptr || sums == nullptr || stats == nullptr)
        return fail(VG_ERR_ARG, "bn_apply_from_sums: null pointer");
    rc = check_channels(dt, C);
    if (rc != VG_OK) return rc;
    const int V = dt == VG_BF16 ? 8 : 4;
    if (C / V > kThreads) return fail(VG_ERR_SHAPE, "bn_apply_from_sums: at most %d channels", kThreads * V);
    if (groups < 1 || groups > 64) return fail(VG_ERR_SHAPE, "bn_apply_from_sums: bad group count %d", groups);
    if (rows < 2) return fail(VG_ERR_SHAPE, "Expected more than 1 value per channel when training");
    const long long nvec = rows * C / V;
    // one resident wave (3 blocks of 256 threads per SM at 80 registers) shared by the groups: every thread pays the
    // per-channel finalisation once and then streams with a grid-stride loop
    const dim3 grid(grid_affine(nvec, C / V, bn_wave_blocks(groups)), groups);
    cudaStream_t st = as_stream(stream);
    if (dt == VG_BF16)
        launch_k(bn_apply_from_sums_kernel<__nv_bfloat16>, dim3(grid), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), nvec, C, rows, sums, gamma, beta,
            running_mean, running_var, num_batches_tracked, momentum, eps, stats, act, slope);
    else
        launch_k(bn_apply_from_sums_kernel<float>, dim3(grid), dim3(kThreads), 0, st, static_cast<const float*>(x), static_cast<float*>(y), nvec, C, rows, sums, gamma, beta, running_mean,
            running_var, num_batches_tracked, momentum, eps, stats, act, slope);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_bn_bwd_apply_from_sums(const void* dz, const void* x, VgDType dt, long long rows, int C, int groups,
                                         const float* stats, const float* sums, float* dgamma, float* dbeta, void* dx,
                                         void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (dz == nullptr || x == nullptr || dx == nullptr || stats == nullptr || sums == nullptr)
        return fail(VG_ERR_ARG, "bn_bwd_apply_from_sums: null pointer");
    rc = check_channels(dt, C);
    if (rc != VG_OK) return rc;
    const int V = dt == VG_BF16 ? 8 : 4;
    if (C / V > kThreads) return fail(VG_ERR_SHAPE, "bn_bwd_apply_from_sums: at most %d channels", kThreads * V);
    if (groups < 1 || groups > 64) return fail(VG_ERR_SHAPE, "bn_bwd_apply_from_sums: bad group count %d", groups);
    const long long nvec = rows * C / V;
    const dim3 grid(grid_affine(nvec, C / V, bn_wave_blocks(groups)), groups);
    cudaStream_t st = as_stream(stream);
    if (dt == VG_BF16)
        launch_k(bn_bwd_apply_from_sums_kernel<__nv_bfloat16>, dim3(grid), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(dz), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(dx),
            nvec, C, rows, stats, sums, dgamma, dbeta);
    else
        launch_k(bn_bwd_apply_from_sums_kernel<float>, dim3(grid), dim3(kThreads), 0, st, static_cast<const float*>(dz), static_cast<const float*>(x), static_cast<float*>(dx), nvec, C, rows, stats,
            sums, dgamma, dbeta);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" size_t vg_bn_bwd_workspace_bytes(long long rows, int channels) {
    (void)rows;
    return 2 * static_cast<size_t>(channels) * sizeof(float);   // the two per-channel projection coefficients
}

extern "C" int vg_act_bwd(const void* dy, const void* x, VgDType in_dt, long long n, VgAct act, float slope, void* dx,
                          VgDType out_dt, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (dy == nullptr || x == nullptr || dx == nullptr) return fail(VG_ERR_ARG, "act_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    const int g = grid_for(n);
    if (in_dt == VG_BF16 && out_dt == VG_BF16)
        launch_k(act_bwd_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(g), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(dx),
            n, act, slope);
    else if (in_dt == VG_F32 && out_dt == VG_BF16)
        launch_k(act_bwd_kernel<float, __nv_bfloat16>, dim3(g), dim3(kThreads), 0, st, static_cast<const float*>(dy),
                                                                     static_cast<const float*>(x),
                                                                     static_cast<__nv_bfloat16*>(dx), n, act, slope);
    else if (in_dt == VG_F32 && out_dt == VG_F32)
        launch_k(act_bwd_kernel<float, float>, dim3(g), dim3(kThreads), 0, st, static_cast<const float*>(dy), static_cast<const float*>(x),
                                                             static_cast<float*>(dx), n, act, slope);
    else
        launch_k(act_bwd_kernel<__nv_bfloat16, float>, dim3(g), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(dy),
                                                                     static_cast<const __nv_bfloat16*>(x),
                                                                     static_cast<float*>(dx), n, act, slope);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_colsum(const void* x, VgDType dt, long long rows, int C, float* out, float* ws, size_t ws_bytes,
                         void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (x == nullptr || out == nullptr) return fail(VG_ERR_ARG, "colsum: null pointer");
    {
        const int V = dt == VG_BF16 ? 8 : 4;
        if (C % V != 0 || !is_pow2(C / V)) {  // odd channel counts (e.g. latent_dim 100): one thread per channel
            if (dt == VG_BF16)
                launch_k(colsum_generic_kernel<__nv_bfloat16>, dim3((C + 127) / 128), dim3(128), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(x), rows, C, out);
            else
                launch_k(colsum_generic_kernel<float>, dim3((C + 127) / 128), dim3(128), 0, as_stream(stream), static_cast<const float*>(x), rows, C, out);
            VG_LAUNCHED();
            return VG_OK;
        }
    }
    (void)ws;
    (void)ws_bytes;
    if (C > kMaxChannels) return fail(VG_ERR_SHAPE, "colsum: more than %d channels", kMaxChannels);
    const ReducePlan p = plan_reduce(rows, std::max(1, kThreads / (C / (dt == VG_BF16 ? 8 : 4))));
    ReduceArgs a{};
    a.x = x; a.rows = rows; a.C = C; a.rows_per_block = p.rows_per_block; a.colsum_out = out;
    return launch_reduce<2>(dt, a, p.blocks, as_stream(stream));
}

extern "C" int vg_nchw_to_nhwc(const float* src, const float* aux, void* dst, VgDType dt, int B, int C, int H, int W,
                               int Cd, int mode, float sigma, int clamp, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (src == nullptr || dst == nullptr || (mode != 0 && aux == nullptr))
        return fail(VG_ERR_ARG, "nchw_to_nhwc: null pointer");
    if (Cd < C) return fail(VG_ERR_SHAPE, "nchw_to_nhwc: dst_channels %d < channels %d", Cd, C);
    const long long HW = static_cast<long long>(H) * W;
    const int g = grid_for(B * HW);
    if (dt == VG_BF16)
        launch_k(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(g), dim3(kThreads), 0, as_stream(stream), src, aux, static_cast<__nv_bfloat16*>(dst), B, C, Cd, HW, mode, sigma, clamp);
    else
        launch_k(nchw_to_nhwc_kernel<float>, dim3(g), dim3(kThreads), 0, as_stream(stream), src, aux, static_cast<float*>(dst), B, C, Cd,
                                                                          HW, mode, sigma, clamp);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_nchw_to_s2d(const float* src, const float* aux, void* dst, int B, int C, int H, int W, int origin,
                              int mode, float sigma, int clamp, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (src == nullptr || dst == nullptr || (mode != 0 && aux == nullptr)) return fail(VG_ERR_ARG, "nchw_to_s2d: null pointer");
    if (C < 1 || C > 16 || (H & 1) || (W & 1) || origin < 0 || origin > 1)
        return fail(VG_ERR_SHAPE, "nchw_to_s2d: needs <= 16 channels, even H and W, origin 0 or 1");
    if (reinterpret_cast<uintptr_t>(dst) & 15) return fail(VG_ERR_ALIGN, "nchw_to_s2d: 16-byte alignment");
    const long long blocks = static_cast<long long>(B) * (H / 2 + origin) * (W / 2 + origin);
    launch_k(nchw_to_s2d_kernel, dim3(grid_for(blocks)), dim3(kThreads), 0, as_stream(stream), src, aux, static_cast<__nv_bfloat16*>(dst), B, C, H, W, origin, mode, sigma, clamp);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_s2d_to_nchw(const void* src, float* dst, int B, int C, int H, int W, int origin, VgAct act, float slope,
                              void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (src == nullptr || dst == nullptr) return fail(VG_ERR_ARG, "s2d_to_nchw: null pointer");
    if (C < 1 || C > 16 || (H & 1) || (W & 1) || origin < 0 || origin > 1)
        return fail(VG_ERR_SHAPE, "s2d_to_nchw: needs <= 16 channels, even H and W, origin 0 or 1");
    if (reinterpret_cast<uintptr_t>(src) & 15) return fail(VG_ERR_ALIGN, "s2d_to_nchw: 16-byte alignment");
    const long long blocks = static_cast<long long>(B) * (H / 2 + origin) * (W / 2 + origin);
    launch_k(s2d_to_nchw_kernel, dim3(grid_for(blocks)), dim3(kThreads), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(src), dst, B, C, H, W, origin, act, slope);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_u8_nhwc_to_nchw(const void* src, float* dst, int B, int C, int H, int W, float mean, float std,
                                  void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (src == nullptr || dst == nullptr) return fail(VG_ERR_ARG, "u8_nhwc_to_nchw: null pointer");
    if (std == 0.f) return fail(VG_ERR_ARG, "u8_nhwc_to_nchw: std must not be zero");
    const long long HW = static_cast<long long>(H) * W;
    launch_k(u8_nhwc_to_nchw_kernel, dim3(grid_for(B * HW)), dim3(kThreads), 0, as_stream(stream), static_cast<const unsigned char*>(src), dst, B, C, HW, mean, 1.f / std);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_linear_permute(const float* src, float* dst, int n_valid, int n_rows, int C, int kk, int mode,
                                 void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (src == nullptr || dst == nullptr || n_valid < 1 || n_rows < n_valid || C < 1 || kk < 1 || (mode != 0 && mode != 1))
        return fail(VG_ERR_ARG, "linear_permute: bad argument");
    const dim3 grid((kk + 31) / 32, (C + 31) / 32, mode == 0 ? n_rows : n_valid);
    launch_k(linear_permute_kernel, dim3(grid), dim3(256), 0, as_stream(stream), src, dst, n_valid, C, kk, mode);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_gather_f32(float* dst, const float* src, const int* idx, long long n, int fan, int accumulate,
                             void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (dst == nullptr || src == nullptr || idx == nullptr || n < 0 || fan < 1)
        return fail(VG_ERR_ARG, "gather_f32: bad argument");
    if (n == 0) return VG_OK;
    launch_k(gather_f32_kernel, dim3(grid_for(n)), dim3(kThreads), 0, as_stream(stream), dst, src, idx, n, fan, accumulate);
    VG_LAUNCHED();
    return VG_OK;
}

extern "C" int vg_nhwc_to_nchw(const void* src, VgDType dt, int Cs, float* dst, int B, int C, int H, int W, VgAct act,
                               float slope, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (src == nullptr || dst == nullptr) return fail(VG_ERR_ARG, "nhwc_to_nchw: null pointer");
    if (Cs < C) return fail(VG_ERR_SHAPE, "nhwc_to_nchw: src_channels %d < channels %d", Cs, C);
    const long long HW = static_cast<long long>(H) * W;
    const int g = grid_for(B * HW);
    if (dt == VG_BF16)
        launch_k(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(g), dim3(kThreads), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(src), dst, B, C, Cs, HW, act, slope);
    else
        launch_k(nhwc_to_nchw_kernel<float>, dim3(g), dim3(kThreads), 0, as_stream(stream), static_cast<const float*>(src), dst, B, C, Cs,
                                                                          HW, act, slope);
    VG_LAUNCHED();
    return VG_OK;
}


/* Fused train-mode BatchNorm + activation forward: one cooperative launch (statistics, running-stat update, affine,
 * activation).  stats = [4][C] (mean, rstd, scale, shift) is written for the backward pass. */
extern "C" int vg_bn_act_train_fwd(const void* x, VgDType dt, long long rows, int C, const float* gamma,
                                   const float* beta, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float momentum, float eps, VgAct act, float slope,
                                   float* stats, void* y, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (x == nullptr || stats == nullptr || y == nullptr) return fail(VG_ERR_ARG, "bn_act_train_fwd: null pointer");
    if (rows < 2) return fail(VG_ERR_SHAPE, "Expected more than 1 value per channel when training");
    FusedArgs a{};
    a.x = x; a.out = y; a.rows = rows; a.C = C; a.act = act; a.slope = slope;
    a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var;
    a.num_batches_tracked = num_batches_tracked; a.momentum = momentum; a.eps = eps; a.stats = stats;
    rc = launch_fused<false>(dt, a, as_stream(stream));
    if (rc != VG_ERR_SHAPE) return rc;
    // shapes the fused kernel does not take: statistics kernel + apply kernel
    rc = vg_bn_train_fwd(x, dt, rows, C, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, stats,
                         stats + C, stats + 2 * C, stats + 3 * C, nullptr, 0, stream);
    if (rc != VG_OK) return rc;
    return vg_scale_shift_act(x, dt, rows, C, stats + 2 * C, stats + 3 * C, act, slope, y, dt, stream);
}

/* Fused backward of act(BN(x)): one cooperative launch; stats as written by the forward. */
extern "C" int vg_bn_act_train_bwd(const void* dy, const void* x, VgDType dt, long long rows, int C, const float* stats,
                                   VgAct act, float slope, float* dgamma, float* dbeta, void* dx, float* ws,
                                   size_t ws_bytes, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (dy == nullptr || x == nullptr || dx == nullptr || stats == nullptr)
        return fail(VG_ERR_ARG, "bn_act_train_bwd: null pointer");
    FusedArgs a{};
    a.x = x; a.dy = dy; a.out = dx; a.rows = rows; a.C = C; a.act = act; a.slope = slope;
    a.stats = const_cast<float*>(stats); a.dgamma = dgamma; a.dbeta = dbeta;
    rc = launch_fused<true>(dt, a, as_stream(stream));
    if (rc != VG_ERR_SHAPE) return rc;
    return vg_bn_act_bwd(dy, x, dt, rows, C, stats + 2 * C, stats + 3 * C, stats, stats + C, act, slope, dgamma, dbeta, dx,
                         ws, ws_bytes, stream);
}
