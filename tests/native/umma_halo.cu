// Probe (test infrastructure): can a K-major SWIZZLE_128B UMMA operand START at an arbitrary 128-byte row of a larger
// shared-memory tile, with 8-row groups a non-multiple-of-1024 stride apart?  That is what "halo tiles" need
// (DESIGN.md section 9): one (th+2) x (tw+2) pixel tile of the activation is loaded once and every shifted tap of a
// stride-1 / phase-decomposed convolution reads its own window of it - window row m of the 128-row operand is pixel
// (m / 8, m % 8) of an 8-wide tile, i.e. halo row  shift + (m / 8) * pitch + (m % 8)  with pitch = tw + 2 = 10.
//
// The tile is written with the address-based 128-byte swizzle TMA uses (16-byte chunk j of the row at byte address a
// goes to chunk j ^ ((a >> 7) & 7)).  Each variant multiplies the window by an identity-like B (N = 64, K = 64) whose
// result makes the rows the tensor core actually read visible, and compares with the expected rows.
//   variants: shift in {0, 1, 3, 8, 11}, pitch in {8 (contiguous), 10 (halo)}, descriptor base-offset field
//   (bits 49..51) = 0 or (start address >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/umma_halo tests/native/umma_halo.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../vae-gan-based-model-for-image-generation-and-denoising_b200/csrc/ptx.cuh"
using namespace vg;

constexpr int kHaloRows = 256;       // rows of 64 bf16 (128 B) in the halo tile
constexpr int kN = 64;

// out[m][n] = sum_k A[row(m)][k] * B[n][k];  A[r][k] = (k == 0 ? r : (k == 1 ? 1 : 0)),  B[n][k] = (k == n % 2)
// -> out[m][n] = n even ? row(m) : 1   (row indices < 256 are exact in bf16)
__global__ void __launch_bounds__(128) halo_kernel(int shift, int pitch, int use_base_offset, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                              // kHaloRows x 128 B
    uint8_t* sB = smem + kHaloRows * 128;            // 64 x 128 B, 1024-aligned
    __shared__ uint64_t done;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kHaloRows * 8; i += blockDim.x) {          // one 16-byte chunk per iteration
        const int r = i / 8, j = i % 8;
        __nv_bfloat16 v[8];
        for (int e = 0; e < 8; ++e) {
            const int k = j * 8 + e;
            v[e] = __float2bfloat16(k == 0 ? static_cast<float>(r) : (k == 1 ? 1.f : 0.f));
        }
        const uint32_t row_addr = smem_u32(sA) + r * 128;
        const int jp = j ^ ((row_addr >> 7) & 7);
        *reinterpret_cast<uint4*>(sA + r * 128 + jp * 16) = *reinterpret_cast<const uint4*>(v);
    }
    for (int i = threadIdx.x; i < kN * 8; i += blockDim.x) {
        const int n = i / 8, j = i % 8;
        __nv_bfloat16 v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((j * 8 + e) == (n & 1) ? 1.f : 0.f);
        const int jp = j ^ (n & 7);
        *reinterpret_cast<uint4*>(sB + n * 128 + jp * 16) = *reinterpret_cast<const uint4*>(v);
    }
    if (threadIdx.x == 0) { mbar_init(&done, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&slot, 64); tmem_relinquish(); }
    fence_proxy_async();                 // generic-proxy writes above -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;

    if (warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, kN, 0, 0);
        const uint32_t a_start = smem_u32(sA) + shift * 128;
        for (int k = 0; k < 4; ++k) {                                        // K = 64 = 4 x 16
            uint64_t ad = make_smem_desc(a_start + k * 32, 0, pitch * 128, 2);
            if (use_base_offset) ad |= static_cast<uint64_t>((a_start >> 7) & 7) << 49;
            const uint64_t bd = make_smem_desc(smem_u32(sB) + k * 32, 0, 1024, 2);
            umma_bf16(tm, ad, bd, idesc, k != 0);
        }
        umma_commit(&done);
    }
    mbar_wait(&done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t taddr = tm + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < kN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[row * kN + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 64);
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 128 * kN * sizeof(float));
    const int smem_bytes = kHaloRows * 128 + kN * 128 + 1024;
    cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    std::vector<float> h(128 * kN);
    int failures = 0;
    for (int pitch : {8, 10})
        for (int shift : {0, 1, 3, 8, 11})
            for (int bo : {0, 1}) {
                cudaMemset(d_out, 0xFF, 128 * kN * sizeof(float));
                halo_kernel<<<1, 128, smem_bytes>>>(shift, pitch, bo, d_out);
                const cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("pitch %2d shift %2d base_offset %d: CUDA error %s\n", pitch, shift, bo, cudaGetErrorString(e));
                    return 2;
                }
                cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
                int bad = 0, first_m = -1;
                float first_got = 0.f;
                for (int m = 0; m < 128; ++m) {
                    const float want = static_cast<float>(shift + (m / 8) * pitch + (m % 8));
                    for (int n = 0; n < kN; ++n) {
                        const float w = (n & 1) ? 1.f : want;
                        if (h[m * kN + n] != w) {
                            if (!bad) { first_m = m; first_got = h[m * kN + (n & ~1)]; }
                            ++bad;
                        }
                    }
                }
                if (bad) {
                    printf("pitch %2d shift %2d base_offset %d: MISMATCH (%d values; first at window row %d: read halo "
                           "row %.0f, wanted %d)\n", pitch, shift, bo, bad, first_m, first_got,
                           shift + (first_m / 8) * pitch + (first_m % 8));
                    // which halo row did each of the first 16 window rows come from?
                    printf("    rows read:");
                    for (int m = 0; m < 16; ++m) printf(" %.0f", h[m * kN]);
                    printf("\n");
                    ++failures;
                } else {
                    printf("pitch %2d shift %2d base_offset %d: ok\n", pitch, shift, bo);
                }
            }
    printf(failures ? "%d variants mismatch\n" : "ALL VARIANTS OK\n", failures);
    return 0;       // a probe, not a gate: the table above is the result
}
