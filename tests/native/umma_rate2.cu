// Micro-benchmark (test infrastructure): raw tcgen05.mma rate for the weight-gradient kernel's operand shapes -
// both operands MN-major SWIZZLE_128B, 128-pixel stages (8 k-steps of 16 rows, atoms LBO = 16 KB apart) - with a
// lean issue loop (descriptors precomputed, only the address word moves).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vae-gan-based-model-for-image-generation-and-denoising_b200/csrc/ptx.cuh"
using namespace vg;

template <int KSTEPS>
__global__ void __launch_bounds__(128) rate_kernel(int n, int mn, int reps, long long* out, int lbo, int kstride16,
                                                   int rotate, int commit) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n, mn, mn);
        const uint64_t ad = make_smem_desc(smem_u32(smem), mn ? lbo : 0, 1024, 2);
        const uint64_t bd = make_smem_desc(smem_u32(smem + 48 * 1024), mn ? lbo : 0, 1024, 2);
        const uint32_t a_lo = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32), b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
        const long long t0 = clock64();
        __shared__ uint64_t cb[4];
        for (int i = 0; i < 4; ++i) mbar_init(&cb[i], 1);
        fence_mbar_init();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tm + (rotate ? (r & 3) * n : 0);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
                umma_bf16_lohi(d, a_lo + k * kstride16, a_hi, b_lo + k * kstride16, b_hi, idesc, 1);
            if (commit) umma_commit(&cb[r & 3]);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(rate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 2000;
    for (int n : {64, 128})
        for (int rotate : {0, 1})
            for (int commit : {0, 1}) {
                rate_kernel<8><<<148, 128, 100 * 1024>>>(n, 1, reps, d, 16384, 2048 >> 4, rotate, commit);
                cudaError_t e = cudaDeviceSynchronize();
                long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
                printf("N=%3d MN/MN 8 k-steps/stage, accumulator %s, commit per stage %d: %7.1f cycles / UMMA (%s)\n", n,
                       rotate ? "rotates over 4" : "fixed", commit, (double)cyc / (reps * 8), cudaGetErrorString(e));
            }
    return 0;
}
