// Data-parallel optimizer step as ONE kernel per gradient bucket over NVLink / NVSwitch peer memory:
//
//     reduce-scatter(gradients)  ->  Adam on this rank's 1/world slice  ->  all-gather(parameters)
//
// The flat gradient and parameter buffers of a network live in symmetric memory (the same virtual layout on every
// rank, peer-mapped, plus one NVSwitch MULTICAST mapping when the fabric offers it).  For a bucket [lo, lo + n) every
// rank owns the slice rank * n / world .. (rank + 1) * n / world:
//   * multicast path: `multimem.ld_reduce.add.v4.f32` pulls the SUM of all replicas' gradients for four elements
//     out of the switch (the reduction happens in the NVSwitch, one 16-byte response per request), the thread applies
//     Adam to its own replica's parameters / moments (the moments of a slice exist on its owner only - the optimizer
//     state is sharded), and `multimem.st.v4.f32` broadcasts the new parameters into every replica;
//   * peer path (no multicast object): the same with `world` peer loads and `world` peer stores per vector.
// Compared with all-reduce + a replicated Adam this moves 1/world of the gradient bytes through each GPU's links
// twice instead of 2 (world - 1) / world times per ring step, runs Adam over 1/world of the state (28 B / parameter
// of HBM traffic / world), needs no NCCL kernel (no ring latency: ~40 us at 8 GPUs for the small tail buckets), and
// keeps the replicas bit-identical by construction (one owner computes each parameter).
// torch.optim.Adam semantics as in losses.cu (vaegan_code.py:42-44,105,134-135); grad_scale = 1 / world.
//
// Cross-GPU barriers: block b of every rank meets block b of every other rank through a signal pad in symmetric
// memory ([block][source rank] epochs, monotonically increasing; release / acquire at system scope).  A barrier that
// does not complete within ~2^26 polls raises *err_flag instead of hanging the device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "common.cuh"
#include "pdl.cuh"

namespace vg {
namespace {

constexpr int kDpThreads = 512;
constexpr int kDpMaxBlocks = 64;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 r;       // (relaxed system-scope load: the data was written by another GPU before the barrier)
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// block `blockIdx.x` of every rank arrives; returns false on timeout
__device__ __forceinline__ bool block_barrier_all_ranks(const VgDpComm& c, unsigned int epoch, int* err_flag) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();            // (also: every thread of this block has issued its preceding global / peer stores)
    if (threadIdx.x < static_cast<unsigned>(c.world)) {
        const int t = threadIdx.x;
        // (a barrier that timed out once poisons the flag: later launches do not wait again)
        const bool poisoned = err_flag != nullptr && *reinterpret_cast<volatile int*>(err_flag) != 0;
        __threadfence_system();
        st_release_sys(c.peer_sig[t] + blockIdx.x * c.world + c.rank, epoch);
        const unsigned int* mine = c.peer_sig[c.rank] + blockIdx.x * c.world + t;
        unsigned int polls = 0;
        while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
            if (poisoned || ++polls > (1u << 25)) {
                s_ok = 0;
                if (err_flag != nullptr) atomicExch(err_flag, 1);
                break;
            }
        }
    }
    __syncthreads();
    return s_ok != 0;
}

template <bool MC>
__global__ void __launch_bounds__(kDpThreads) dp_adam_bucket_kernel(const VgDpComm c, long long lo, long long n,
                                                                   float* __restrict__ m, float* __restrict__ v,
                                                                   double lr_d, double b1_d, double b2_d, double eps_d,
                                                                   const long long* __restrict__ step_ptr,
                                                                   float grad_scale, int write_grads, int* err_flag) {
    pdl_enter();
    // two barriers per launch: epochs base + 1 and base + 2; the block that finishes last advances the counter
    const unsigned int base = *reinterpret_cast<volatile unsigned int*>(c.epoch);
    const double t = static_cast<double>(*step_ptr);
    const float b2 = static_cast<float>(b2_d), eps = static_cast<float>(eps_d);
    const float omb1 = static_cast<float>(1.0 - b1_d), omb2 = static_cast<float>(1.0 - b2_d);
    const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(b2_d, t)));
    const float step_size = static_cast<float>(lr_d / (1.0 - pow(b1_d, t)));

    // every rank's gradients of this bucket are complete (they were written by kernels that precede this launch in
    // stream order on their own GPU; the barrier carries that across GPUs)
    bool ok = block_barrier_all_ranks(c, base + 1, err_flag);

    const long long nvec = n / 4;
    const long long per = (nvec + c.world - 1) / c.world;
    const long long v_begin = per * c.rank, v_end = min(nvec, per * (c.rank + 1));
    float* p_loc = c.peer_params[c.rank];
    constexpr int U = 4;        // independent 16-byte requests per thread in flight (NVLink round trips are ~2-3 us)
    if (ok) {
        const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
        for (long long i0 = v_begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < v_end;
             i0 += stride * U) {
            float4 g[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * stride;
                if (i >= v_end) continue;
                const long long e = lo + i * 4;
                if (MC) {
                    g[u] = multimem_ld_reduce_add(c.mc_grads + e);
                } else {
                    g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r = 0; r < c.world; ++r) {
                        const float4 q = ld_peer(c.peer_grads[r] + e);
                        g[u].x += q.x; g[u].y += q.y; g[u].z += q.z; g[u].w += q.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * stride;
                if (i >= v_end) continue;
                const long long e = lo + i * 4;
                if (write_grads) {              // parity tests: the summed gradient, visible on every rank afterwards
                    if (MC) multimem_st(c.mc_grads + e, g[u]);
                    else
                        for (int r = 0; r < c.world; ++r) st_peer(c.peer_grads[r] + e, g[u]);
                }
                const float4 pv = *reinterpret_cast<const float4*>(p_loc + e);
                const float4 mv = *reinterpret_cast<const float4*>(m + e), vv = *reinterpret_cast<const float4*>(v + e);
                float pp[4] = {pv.x, pv.y, pv.z, pv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w};
                float vq[4] = {vv.x, vv.y, vv.z, vv.w};
                const float gg[4] = {g[u].x * grad_scale, g[u].y * grad_scale, g[u].z * grad_scale, g[u].w * grad_scale};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mm[j] = mm[j] + (gg[j] - mm[j]) * omb1;            // exp_avg.lerp_(grad, 1-beta1)
                    vq[j] = vq[j] * b2 + omb2 * gg[j] * gg[j];         // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
                    const float denom = sqrtf(vq[j]) / bc2_sqrt + eps;
                    pp[j] = pp[j] - step_size * (mm[j] / denom);
                }
                *reinterpret_cast<float4*>(m + e) = make_float4(mm[0], mm[1], mm[2], mm[3]);
                *reinterpret_cast<float4*>(v + e) = make_float4(vq[0], vq[1], vq[2], vq[3]);
                const float4 pn = make_float4(pp[0], pp[1], pp[2], pp[3]);
                if (MC) multimem_st(c.mc_params + e, pn);
                else
                    for (int r = 0; r < c.world; ++r) st_peer(c.peer_params[r] + e, pn);
            }
        }
    }
    // every slice of the new parameters has landed in every replica before any rank's next kernel reads them
    block_barrier_all_ranks(c, base + 2, err_flag);
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(c.epoch + 1, 1u) == gridDim.x - 1) {
            c.epoch[1] = 0;
            *reinterpret_cast<volatile unsigned int*>(c.epoch) = base + 2;
        }
    }
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" int vg_dp_max_blocks(void) { return kDpMaxBlocks; }

extern "C" int vg_dp_adam_bucket(const VgDpComm* comm, long long lo, long long n, float* m, float* v, double lr,
                                 double beta1, double beta2, double eps, const long long* step_dev, float grad_scale,
                                 int blocks, int write_grads, int* err_flag, void* stream) {
    int rc = device_check();
    if (rc != VG_OK) return rc;
    if (comm == nullptr || m == nullptr || v == nullptr || step_dev == nullptr)
        return fail(VG_ERR_ARG, "dp_adam_bucket: null pointer");
    const VgDpComm& c = *comm;
    if (c.world < 2 || c.world > VG_DP_MAX_RANKS || c.rank < 0 || c.rank >= c.world)
        return fail(VG_ERR_ARG, "dp_adam_bucket: bad rank / world (%d / %d)", c.rank, c.world);
    if (c.epoch == nullptr) return fail(VG_ERR_ARG, "dp_adam_bucket: null epoch counter");
    for (int r = 0; r < c.world; ++r)
        if (c.peer_grads[r] == nullptr || c.peer_params[r] == nullptr || c.peer_sig[r] == nullptr)
            return fail(VG_ERR_ARG, "dp_adam_bucket: missing peer mapping of rank %d", r);
    if ((c.mc_grads == nullptr) != (c.mc_params == nullptr))
        return fail(VG_ERR_ARG, "dp_adam_bucket: multicast mappings must come as a pair");
    if ((lo | n) & 3) return fail(VG_ERR_ALIGN, "dp_adam_bucket: bucket bounds must be multiples of 4 elements");
    if (n <= 0) return VG_OK;
    if (blocks < 1 || blocks > kDpMaxBlocks) return fail(VG_ERR_ARG, "dp_adam_bucket: 1..%d blocks", kDpMaxBlocks);
    cudaError_t e;
    if (c.mc_grads != nullptr)
        e = launch_k(dp_adam_bucket_kernel<true>, dim3(blocks), dim3(kDpThreads), 0, as_stream(stream), c, lo, n, m, v,
                     lr, beta1, beta2, eps, step_dev, grad_scale, write_grads, err_flag);
    else
        e = launch_k(dp_adam_bucket_kernel<false>, dim3(blocks), dim3(kDpThreads), 0, as_stream(stream), c, lo, n, m, v,
                     lr, beta1, beta2, eps, step_dev, grad_scale, write_grads, err_flag);
    if (e != cudaSuccess) return cuda_fail(e, "dp_adam_bucket launch");
    VG_LAUNCHED();
    return VG_OK;
}
