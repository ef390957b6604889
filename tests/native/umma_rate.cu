// Micro-benchmark (test infrastructure): issue rate of tcgen05.mma M=128 for K-major vs MN-major operands and
// several N, on whatever bytes are in shared memory.  One CTA per SM, one issuing thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vae-gan-based-model-for-image-generation-and-denoising_b200/csrc/ptx.cuh"
using namespace vg;

__global__ void __launch_bounds__(128) rate_kernel(int n, int a_mn, int b_mn, int reps, long long* out, int lbo, int sbo_a, int sbo_b, int mode, int tcols) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&slot, tcols); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n, a_mn, b_mn);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = a_mn ? make_smem_desc(a_addr + k * 2048, lbo, 1024, 2) : make_smem_desc(a_addr + k * 32, 0, sbo_a, 2);
                const uint64_t bd = b_mn ? make_smem_desc(b_addr + k * 2048, lbo, 1024, 2) : make_smem_desc(b_addr + k * 32, 0, sbo_b, 2);
                umma_bf16(tm + ((mode & 2) ? (r & 3) * n : 0), ad, bd, idesc, (mode & 4) ? (k != 0 || r >= 4) : 1);
            }
            if (mode & 1) umma_commit(&bar2[r & 7]);
            if (mode & 8) { mbar_wait(&bar2[r & 7], (r >> 3) & 1); tc_fence_after(); }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tm, tcols);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 2000;
    for (int tcols : {128, 256, 512})
        for (int n : {64, 128})
            for (int mn : {0, 1})
                for (int mode : {0, 2}) {
                    if (mode == 2 && 4 * n > tcols) continue;
                    rate_kernel<<<148, 128, 100 * 1024>>>(n, mn, mn, reps, d, 64 * 128, 1024, 1024, mode, tcols);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
                    printf("tmem cols %3d N=%3d %s mode %d: %7.1f cycles / UMMA  (%s)\n", tcols, n, mn ? "MN/MN" : "K/K  ", mode,
                           (double)cyc / (reps * 4), cudaGetErrorString(e));
                }
    return 0;
}
