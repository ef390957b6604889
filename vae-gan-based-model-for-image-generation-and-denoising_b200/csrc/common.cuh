// Shared host-side helpers for the C-ABI layer: error reporting and device checks.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/vaegan_b200.h"

namespace vg {

char* last_error_buffer();  // thread-local, 512 bytes

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

inline int cuda_fail(cudaError_t e, const char* what) {
    return fail(VG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Use right after a kernel launch: checks the launch and counts it.
#define VG_LAUNCHED()                                          \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return ::vg::cuda_fail(e__, "kernel launch"); \
        ::vg::note_launch();                                   \
    } while (0)

#define VG_CUDA(call)                                          \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ::vg::cuda_fail(e__, #call); \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// sm_100 check, cached per device.
int device_check();

// Number of kernels this library has launched in this process (bench.py reports it as gpu_launches).
void note_launch(int n = 1);

}  // namespace vg
