// 16-byte vector access for the HBM-bound kernels: 4 floats or 8 bf16 per load / store, arithmetic in fp32.
#pragma once
#include <cuda_bf16.h>

#include <cstdint>

namespace vg {

template <typename T>
struct Vec;
template <>
struct Vec<float> {
    static constexpr int N = 4;
    using Raw = float4;
    __device__ static Raw load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
    __device__ static void unpack(const Raw& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ static void load(const float* p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    // (a batch of loads in flight is kept PACKED - 4 registers per 16 bytes instead of 8 floats - and unpacked at use)
    using Raw = uint4;
    __device__ static Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
    __device__ static void unpack(const Raw& t, float (&v)[8]) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

}  // namespace vg
