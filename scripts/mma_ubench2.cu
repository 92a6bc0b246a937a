// tcgen05.mma kind::tf32 issue-rate floor: straight-line unrolled MMAs with precomputed descriptors
// (no per-MMA address arithmetic), one issuing thread per CTA, one CTA per SM.  Development tool.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

template <int N, int COMMIT>
__global__ void __launch_bounds__(128, 1) ubench(int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, scratch;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&scratch)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ad[j] = make_desc(a0 + j * 4160, 2080, 128);
            bd[j] = make_desc(b0 + j * (N / 32) * 1024, (N / 32) * 512, 128);
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 12) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma(tmem_base + r * 32, ad[j], bd[(j + r) & 3], idesc);
            if (COMMIT)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&scratch)) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t2 = clock64();
        out[blockIdx.x] = t2 - t0;
        out[148 + blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int N, int COMMIT>
void run(long long* d_out) {
    const int iters = 3600;
    cudaFuncSetAttribute(ubench<N, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) ubench<N, COMMIT><<<148, 128, 200 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return; }
    std::vector<long long> t(296);
    cudaMemcpy(t.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
    std::sort(t.begin(), t.begin() + 148); std::sort(t.begin() + 148, t.end());
    const double cyc = (double)t[74] / iters, iss = (double)t[148 + 74] / iters;
    const double flop = 2.0 * 128 * N * 8;
    printf("N=%3d commit/12=%d: %.1f cyc/MMA (issue %.1f) -> %.0f TFLOP/s; model 32+N/4 = %d, compute N/2 = %d\n",
           N, COMMIT, cyc, iss, flop / cyc * 148 * 1.965e9 / 1e12, 32 + N / 4, N / 2);
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 296 * 8);
    run<32, 0>(d_out); run<32, 1>(d_out); run<64, 0>(d_out); run<96, 0>(d_out); run<96, 1>(d_out);
    run<128, 0>(d_out); run<160, 0>(d_out); run<224, 0>(d_out); run<224, 1>(d_out); run<256, 0>(d_out);
    return 0;
}
