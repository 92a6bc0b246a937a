// tcgen05.mma kind::tf32 throughput microbenchmark (development tool, not part of the library).
// One CTA per SM; one thread issues `iters` MMAs back to back, commits, waits; cycles per MMA are
// reported for a list of (M, N, operand layout, accumulator rotation) cases.  The operand layouts
// are the ones the conv engine uses or could use: no-swizzle K-major core matrices with a plane
// pitch as LBO ("quad planes"), and the canonical 128-byte-swizzled K-major GEMM layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ubench scripts/mma_ubench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

struct Case {
    const char* name;
    int M, N;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int layout;          // 0 = no swizzle, 2 = 128B swizzle
    int nd;              // accumulators rotated over
    uint32_t a_step, a_wrap, b_step, b_wrap;   // bytes added to the start addresses per MMA (mod wrap)
    int iters;
    int nissue;          // warps issuing concurrently (each on its own accumulators)
    uint32_t a_off;      // byte offset of the A start address (operand alignment experiments)
    int commit_every;    // tcgen05.commit to a scratch mbarrier after this many MMAs (0 = only at the end)
    int d_step;          // TMEM column step between consecutive MMAs' accumulators when nd > 1 (0 = N)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128, 1) ubench(Case c, long long* out_total, long long* out_issue) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint64_t scratch_bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1u) : "memory");
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&scratch_bar[i])), "r"(1u) : "memory");
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
    if ((tid & 31) == 0 && warp < c.nissue) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
        const uint32_t a0 = smem_u32(smem) + warp * 8192 + c.a_off, b0 = smem_u32(smem) + 100 * 1024 + warp * 2048;
        uint32_t ao = 0, bo = 0;
        int d = 0, since_commit = 0;
        const long long t0 = clock64();
        for (int i = 0; i < c.iters; ++i) {
            const uint64_t ad = make_desc(a0 + ao, c.a_lbo, c.a_sbo, c.layout);
            const uint64_t bd = make_desc(b0 + bo, c.b_lbo, c.b_sbo, c.layout);
            const uint32_t dt = tmem_base + (c.d_step ? d * c.d_step : (warp * c.nd + d) * c.N);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(dt), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            ao += c.a_step; if (ao >= c.a_wrap) ao = 0;
            bo += c.b_step; if (bo >= c.b_wrap) bo = 0;
            if (++d == c.nd) d = 0;
            if (c.commit_every && ++since_commit == c.commit_every) {
                since_commit = 0;
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&scratch_bar[warp])) : "memory");
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0u) : "memory");
        }
        const long long t2 = clock64();
        if (warp == 0) { out_total[blockIdx.x] = t2 - t0; out_issue[blockIdx.x] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main() {
    const int IT = 4000;
    const uint32_t P = 130 * 16;      // quad-plane pitch of a 128+2 pixel row
    std::vector<Case> cases = {
        // pixels as M (current engine): A = 128 px x 8 k (two quad planes), B = 32 cout x 8 k
        {"M128 N32 px-as-M  same D          ", 128, 32, P, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N32 px-as-M  3 rotating D    ", 128, 32, P, 128, 512, 128, 0, 3, 0, 1, 0, 1, IT, 1},
        {"M128 N32 px-as-M  walk A/B, 3 D   ", 128, 32, P, 128, 512, 128, 0, 3, 2 * P, 8 * P, 1024, 36 * 1024, IT, 1},
        {"M128 N32 px-as-M  pitch 2048      ", 128, 32, 2048, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N64 px-as-M                  ", 128, 64, P, 128, 1024, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N128 px-as-M                 ", 128, 128, P, 128, 2048, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        // weights as M (4 taps x 32 cout), pixels as N: A = 128 x 8 k, B = N px x 8 k (quad planes)
        {"M128 N64  w-as-M                  ", 128, 64, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N128 w-as-M                  ", 128, 128, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N256 w-as-M                  ", 128, 256, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT / 2, 1},
        {"M128 N256 w-as-M walk A/B         ", 128, 256, 2048, 128, 258 * 16, 128, 0, 1, 4096, 36 * 1024, 2 * 258 * 16, 8 * 258 * 16, IT / 2, 1},
        {"M128 N256 w-as-M 2 rotating D     ", 128, 256, 2048, 128, 258 * 16, 128, 0, 2, 4096, 36 * 1024, 2 * 258 * 16, 8 * 258 * 16, IT / 2, 1},
        {"M64  N256 w-as-M                  ", 64, 256, 1024, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT / 2, 1},
        // canonical GEMM layout: 128B swizzle, K-major, SBO = 1024
        {"M128 N32  sw128                   ", 128, 32, 16, 1024, 16, 1024, 2, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N256 sw128                   ", 128, 256, 16, 1024, 16, 1024, 2, 1, 0, 1, 0, 1, IT / 2, 1},
        {"M128 N32  sw128 walk K (32 B)     ", 128, 32, 16, 1024, 16, 1024, 2, 1, 32, 128, 32, 128, IT, 1},
        {"M128 N32 px-as-M  2 issuing warps ", 128, 32, P, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 2},
        {"M128 N32 px-as-M  4 issuing warps ", 128, 32, P, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4},
        {"M128 N64 px-as-M  2 issuing warps ", 128, 64, P, 128, 1024, 128, 0, 1, 0, 1, 0, 1, IT, 2},
        {"M128 N128 w-as-M  2 issuing warps ", 128, 128, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT, 2},
        {"M128 N128 w-as-M  4 issuing warps ", 128, 128, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT, 4},
        {"M128 N192 w-as-M                  ", 128, 192, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"M128 N160 w-as-M                  ", 128, 160, 2048, 128, 258 * 16, 128, 0, 1, 0, 1, 0, 1, IT, 1},
        {"N32 4 issuers: pitch 2048, A aligned     ", 128, 32, 2048, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0},
        {"N32 4 issuers: pitch 2048, A +16 B       ", 128, 32, 2048, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 16},
        {"N32 4 issuers: pitch 2048, A +32 B       ", 128, 32, 2048, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 32},
        {"N32 4 issuers: pitch 2048, A +64 B       ", 128, 32, 2048, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 64},
        {"N32 4 issuers: pitch 2080, A aligned     ", 128, 32, 2080, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0},
        {"N32 4 issuers: pitch 2112, A aligned     ", 128, 32, 2112, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0},
        {"N32 4 issuers: pitch 2176, A aligned     ", 128, 32, 2176, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0},
        {"N32 4 issuers: pitch 2304(+256), aligned ", 128, 32, 2304, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0},
        {"N32 3 issuers: pitch 2048, A aligned     ", 128, 32, 2048, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 3, 0},
        {"N32 4 issuers: walk A (+4160/MMA)        ", 128, 32, 2080, 128, 512, 128, 0, 1, 4160, 4 * 4160, 1024, 36 * 1024, IT, 4, 0},
        {"N64 4 issuers: pitch 2048 aligned        ", 128, 64, 2048, 128, 1024, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0},
        {"N96 conv-engine layout (1 issuer)          ", 128, 96, 2080, 128, 1536, 128, 0, 1, 0, 1, 0, 1, IT, 1, 0, 0, 0},
        {"N96 conv layout, D slides by 32 cols        ", 128, 96, 2080, 128, 1536, 128, 0, 13, 0, 1, 0, 1, IT, 1, 0, 0, 32},
        {"N96 conv layout, commit every 12            ", 128, 96, 2080, 128, 1536, 128, 0, 1, 0, 1, 0, 1, IT, 1, 0, 12, 0},
        {"N96 conv layout, commit every 12, walk A/B  ", 128, 96, 2080, 128, 1536, 128, 0, 13, 4160, 4 * 4160, 3072, 36 * 1024, IT, 1, 0, 12, 32},
        {"N32 conv layout, commit every 12 (1 issuer) ", 128, 32, 2080, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 1, 0, 12, 0},
        {"N32 conv layout, commit every 4  (1 issuer) ", 128, 32, 2080, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 1, 0, 4, 0},
        {"N32 conv layout, commit every 12 (4 issuers)", 128, 32, 2080, 128, 512, 128, 0, 1, 0, 1, 0, 1, IT, 4, 0, 12, 0},
        {"N224 conv layout (k7), commit every 14      ", 128, 224, 2080, 128, 3584, 128, 0, 1, 0, 1, 0, 1, IT, 1, 0, 14, 0},
        {"N96 2 issuers                                ", 128, 96, 2080, 128, 1536, 128, 0, 1, 0, 1, 0, 1, IT, 2, 0, 0, 0},
        {"N64 conv layout (1 issuer)                   ", 128, 64, 2080, 128, 1024, 128, 0, 1, 0, 1, 0, 1, IT, 1, 0, 0, 0},
    };
    long long *d_tot, *d_iss;
    cudaMalloc(&d_tot, 148 * 8); cudaMalloc(&d_iss, 148 * 8);
    cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock (attr) %d kHz\n", clk);
    for (auto& c : cases) {
        for (int rep = 0; rep < 2; ++rep) ubench<<<148, 128, 200 * 1024>>>(c, d_tot, d_iss);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<long long> t(148), s(148);
        cudaMemcpy(t.data(), d_tot, 148 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(s.data(), d_iss, 148 * 8, cudaMemcpyDeviceToHost);
        std::sort(t.begin(), t.end()); std::sort(s.begin(), s.end());
        const double cyc = (double)t[74] / c.iters, iss = (double)s[74] / c.iters;
        const double flop = 2.0 * c.M * c.N * 8 * c.nissue;
        printf("%s cyc/MMA median %.1f (min %.1f max %.1f)  issue %.1f  -> %.0f FLOP/cyc/SM = %.0f TFLOP/s @1.965GHz x148\n",
               c.name, cyc, (double)t[0] / c.iters, (double)t[147] / c.iters, iss, flop / cyc, flop / cyc * 148 * 1.965e9 / 1e12);
    }
    return 0;
}
