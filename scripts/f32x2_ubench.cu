// FP32 scalar vs packed f32x2 (add.rn.f32x2 / fma.rn.f32x2) issue rate on sm_100a, at the occupancy the
// guided-filter kernels run at (6 warps per SM).  Development tool.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float addv(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE, int ILP>
__global__ void k(float* out, int iters, long long* cyc) {
    float s[ILP * 2];
    u64 p[ILP];
    for (int i = 0; i < ILP * 2; ++i) s[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < ILP; ++i) p[i] = ((u64)__float_as_uint(s[2 * i]) << 32) | __float_as_uint(s[2 * i + 1]);
    const float inc = 1.0f + blockIdx.x;
    const u64 pinc = ((u64)__float_as_uint(inc) << 32) | __float_as_uint(inc);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < ILP * 2; ++i) s[i] = addv(s[i], inc);
        } else {
#pragma unroll
            for (int i = 0; i < ILP; ++i) p[i] = add2(p[i], pinc);
        }
    }
    long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < ILP * 2; ++i) r += s[i];
    for (int i = 0; i < ILP; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int threads) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    k<MODE, ILP><<<148, threads>>>(out, iters, cyc);
    k<MODE, ILP><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / iters;
    printf("%-28s threads/SM %4d ILP %2d: %.1f cycles per iteration of %d fp32 adds per lane -> %.2f adds/cycle/lane-group, %.1f adds/clk/SM\n",
           name, threads, ILP, c, ILP * 2, ILP * 2 / c, ILP * 2.0 * threads / c);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 8>("scalar add.f32", 192); run<1, 8>("packed add.f32x2", 192);
    run<0, 8>("scalar add.f32", 512); run<1, 8>("packed add.f32x2", 512);
    run<0, 2>("scalar add.f32", 192); run<1, 2>("packed add.f32x2", 192);
    run<0, 16>("scalar add.f32", 1024); run<1, 16>("packed add.f32x2", 1024);
    return 0;
}
