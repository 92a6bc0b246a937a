// Shared helpers for the paif_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/paif_b200.h"

namespace paif {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define PAIF_REQUIRE(cond, msg)                                      \
    do {                                                             \
        if (!(cond)) {                                               \
            paif::set_error("%s: %s", __func__, msg);                \
            return PAIF_EINVAL;                                      \
        }                                                            \
    } while (0)

// C4 map addressing: [B][C/4][H][W][4]
__host__ __device__ inline size_t c4_plane(int b, int q, int Q, int H, int W) {
    return ((size_t)b * Q + q) * (size_t)H * W;   // in float4 units
}

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : v * a; }
__device__ __forceinline__ float dprelu_f(float src, float a) { return src > 0.f ? 1.f : a; }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) {
    return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}

// C8 bf16 maps ([B][C/8][H][W][8], the bf16 storage mode): a pixel's 8 channels are one 16-byte vector
__device__ __forceinline__ void bf8_unpack(uint4 u, float4& lo, float4& hi) {
    lo = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    hi = make_float4(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u),
                     __uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t bf2_pack(float e0, float e1) {      // round to nearest even; e0 at the lower address
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
__device__ __forceinline__ uint4 bf8_pack(float4 lo, float4 hi) {
    return make_uint4(bf2_pack(lo.x, lo.y), bf2_pack(lo.z, lo.w), bf2_pack(hi.x, hi.y), bf2_pack(hi.z, hi.w));
}

// all 32 channels of one pixel of image b: `img` points at the image's first plane (fp32 C4: 8 planes of float4,
// bf16 C8: 4 planes of uint4), `plane` = H*W, `pix` = y*W + x
template <bool BF>
__device__ __forceinline__ void ld_px32(const void* img, size_t plane, size_t pix, float4 (&v)[8]) {
    if constexpr (BF) {
#pragma unroll
        for (int p = 0; p < 4; ++p) bf8_unpack(__ldg(static_cast<const uint4*>(img) + p * plane + pix), v[2 * p], v[2 * p + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(static_cast<const float4*>(img) + q * plane + pix);
    }
}
template <bool BF>
__device__ __forceinline__ void st_px32(void* img, size_t plane, size_t pix, const float4 (&v)[8]) {
    if constexpr (BF) {
#pragma unroll
        for (int p = 0; p < 4; ++p) static_cast<uint4*>(img)[p * plane + pix] = bf8_pack(v[2 * p], v[2 * p + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) static_cast<float4*>(img)[q * plane + pix] = v[q];
    }
}
// start of image b of a 32-channel map
template <bool BF>
__device__ __forceinline__ const void* img32(const void* map, int b, size_t plane) {
    return static_cast<const uint4*>(map) + (size_t)b * (BF ? 4 : 8) * plane;      // both layouts: 16-byte pixel vectors
}
template <bool BF>
__device__ __forceinline__ void* img32(void* map, int b, size_t plane) {
    return static_cast<uint4*>(map) + (size_t)b * (BF ? 4 : 8) * plane;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per device: `done` holds one bit per device ordinal (a process normally drives one GPU,
// but nothing here assumes it).  Returns true when the attribute still has to be set on the current device.
inline bool attr_needed(unsigned long long& done, int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev >= 64 || !((done >> dev) & 1ull);
}
inline void attr_mark(unsigned long long& done, int dev) { if (dev < 64) done |= 1ull << dev; }

}  // namespace paif
