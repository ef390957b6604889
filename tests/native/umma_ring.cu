// Micro-benchmark (test infrastructure): the weight-gradient kernel's MMA loop in isolation - `ksteps` UMMAs per
// stage into one accumulator, tcgen05.commit per stage, and the stage is re-issued only after the commit of `ring`
// stages earlier has arrived (what the producer/empty-barrier turnaround imposes).  Operands are whatever bytes
// are in shared memory; one CTA per SM, one issuing thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vae-gan-based-model-for-image-generation-and-denoising_b200/csrc/ptx.cuh"
using namespace vg;

__global__ void __launch_bounds__(192) ring_kernel(int n, int mn, int ksteps, int lbo, int ring, int stages_total,
                                                   int accs, long long* out, int commit_every, int do_wait, int fence) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bars[32];
    __shared__ uint64_t done;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(&bars[i], 1); mbar_init(&done, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n, mn, mn);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
        const int stage_bytes = 32 * 1024;
        const long long t0 = clock64();
        for (int s = 0; s < stages_total; ++s) {
            const int slot_i = s % ring;
            if (do_wait && commit_every == 1 && s >= ring) {
                mbar_wait(&bars[slot_i], ((s / ring) - 1) & 1);
                if (fence) tc_fence_after();
            }
            const uint32_t a_addr = a_base + (s % 2) * stage_bytes, b_addr = b_base + (slot_i % 4) * stage_bytes;
            const uint32_t d = tm + (s % accs) * n;
            for (int k = 0; k < ksteps; ++k) {
                const uint64_t ad = mn ? make_smem_desc(a_addr + k * 2048, lbo, 1024, 2) : make_smem_desc(a_addr + k * 32, 0, 1024, 2);
                const uint64_t bd = mn ? make_smem_desc(b_addr + k * 2048, lbo, 1024, 2) : make_smem_desc(b_addr + k * 32, 0, 1024, 2);
                umma_bf16(d, ad, bd, idesc, 1);
            }
            if (commit_every == 1 || (s % commit_every) == commit_every - 1) umma_commit(&bars[slot_i]);
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int stages = 4000;
    struct V { int commit_every, do_wait, fence; const char* what; };
    const V variants[] = {{1, 1, 1, "commit+wait+fence per stage"}, {1, 1, 0, "commit+wait, no fence"},
                          {1, 0, 0, "commit per stage, never wait"}, {4, 0, 0, "commit every 4 stages"},
                          {100000, 0, 0, "no commits"}};
    for (const V& v : variants)
        for (int ksteps : {4, 8}) {
            const int n = 128, mn = 1, lbo = ksteps == 8 ? 16384 : 8192, ring = 8, accs = 4;
            ring_kernel<<<148, 192, 200 * 1024>>>(n, mn, ksteps, lbo, ring, stages, accs, d, v.commit_every, v.do_wait, v.fence);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
            printf("%-32s ksteps %d: %7.1f cycles / UMMA, %8.1f / stage  (%s)\n", v.what, ksteps,
                   (double)cyc / (stages * ksteps), (double)cyc / stages, cudaGetErrorString(e));
        }
    return 0;
}
