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
