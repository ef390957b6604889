// Programmatic dependent launch (PDL): every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel of a stream (or of a captured graph branch)
// is scheduled while its predecessor is still running; its CTAs set themselves up (barrier init, TMEM allocation,
// descriptor prefetch, index arithmetic) and then block in griddepcontrol.wait until the predecessor grid has
// COMPLETED and its memory is visible.  Rules every kernel follows:
//   * pdl_enter() (= launch_dependents + wait) is executed by ALL threads before the first access to global memory;
//     kernels that allocate tensor memory call it after the allocation (a dependent CTA that grabbed TMEM columns
//     first would starve a not-yet-allocated CTA of the grid it is waiting for);
//   * nothing before it touches memory another kernel may write, and nothing at all is written before it.
// Because every kernel waits for its predecessor's completion before it can itself complete, stream order stays
// transitive: a kernel that depends on a grid two launches back still sees its results.
// VG_PDL=0 in the environment launches everything without the attribute (the instructions are then no-ops).
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace vg {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_trigger();
    pdl_wait();
}

bool pdl_enabled();   // common.cu (environment switch VG_PDL, read once)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

}  // namespace vg
