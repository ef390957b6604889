// Probe: are thread-block clusters launchable on this box, and with how much dynamic shared memory?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(1, 2, 1) k_y(int* out) {
    extern __shared__ unsigned char sm[];
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    if (threadIdx.x == 0) { atomicAdd(out, 1 + (int)r * 100); }
}
__global__ void __cluster_dims__(2, 1, 1) k_x(int* out) {
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0) { atomicAdd(out, 1); }
}
int main() {
    int v = -1;
    cudaDeviceGetAttribute(&v, cudaDevAttrClusterLaunch, 0);
    printf("cudaDevAttrClusterLaunch = %d\n", v);
    int* d; cudaMalloc(&d, 4);
    for (int smem : {0, 48 * 1024, 100 * 1024, 198448, 227 * 1024}) {
        cudaFuncSetAttribute(k_y, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_x, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaMemset(d, 0, 4);
        k_y<<<dim3(1, 2, 4), 192, smem>>>(d);
        cudaError_t e1 = cudaGetLastError();
        cudaError_t s1 = cudaDeviceSynchronize();
        k_x<<<dim3(2, 1, 4), 192, smem>>>(d);
        cudaError_t e2 = cudaGetLastError();
        cudaError_t s2 = cudaDeviceSynchronize();
        int h = 0; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
        printf("smem %6d: y-cluster %s / %s ; x-cluster %s / %s ; sum %d\n", smem, cudaGetErrorString(e1),
               cudaGetErrorString(s1), cudaGetErrorString(e2), cudaGetErrorString(s2), h);
    }
    return 0;
}
