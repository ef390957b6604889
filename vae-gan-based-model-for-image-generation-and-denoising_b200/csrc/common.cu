#include "common.cuh"
#include "pdl.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace vg {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = [] {
        const char* v = getenv("VG_PDL");
        return v == nullptr || v[0] != '0';
    }();
    return on;
}

int device_check() {
    static std::mutex mu;
    static int cached[64];
    static bool known[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && known[dev]) {
        if (cached[dev] != VG_OK) fail(cached[dev], "device %d is not sm_100 (this library has no fallback path)", dev);
        return cached[dev];
    }
    int major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    const int rc = (major == 10 && minor == 0) ? VG_OK : VG_ERR_ARCH;
    if (dev >= 0 && dev < 64) {
        cached[dev] = rc;
        known[dev] = true;
    }
    if (rc != VG_OK)
        return fail(rc, "device %d is sm_%d%d, need sm_100 (this library has no fallback path)", dev, major, minor);
    return rc;
}

}  // namespace vg

namespace vg { long long launch_count(); }

extern "C" const char* vg_last_error(void) { return vg::last_error_buffer(); }
extern "C" int vg_version(void) { return 100; }
extern "C" long long vg_launch_count(void) { return vg::launch_count(); }
extern "C" int vg_device_check(void) { return vg::device_check(); }
