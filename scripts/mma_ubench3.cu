// Why does an N=96 tcgen05.mma cost ~110 cycles inside conv_tc_kernel when the straight-line microbenchmark
// (mma_ubench2.cu) reaches 56?  Same issue pattern as the kernel's 3x3 row (12 MMAs: dx 0..2 x k8 0..3, one accumulator
// range), then one factor at a time: tap-shift alignment, accumulator rotation, concurrent shared-memory writes
// (bulk copies like the producer's), concurrent TMEM traffic (tcgen05.ld/st like the epilogue's).  Development tool.
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
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int RW = 130, N = 96;
// MODE bit 0: kernel-like descriptors (else ubench2-like)   bit 1: 32-byte aligned tap shifts
//      bit 2: rotate the accumulator per row                 bit 3: concurrent bulk copies into shared memory
//      bit 4: concurrent tcgen05.ld/st on other TMEM columns bit 5: commit after every row
template <int MODE>
__global__ void __launch_bounds__(160, 1) ubench(int rows, long long* out, const float* gsrc, int delay = 0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, scratch, cbar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 160) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (tid == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&scratch)), "r"(1u) : "memory");
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&cbar[i])), "r"(1u) : "memory");
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
        uint64_t ad[12], bd[12];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
                const int j = dx * 4 + k8;
                if (MODE & 1) {
                    const int shift = (MODE & 2) ? dx * 2 : dx;
                    ad[j] = make_desc(a0 + (shift + k8 * 2 * RW) * 16, RW * 16, 128);
                    bd[j] = make_desc(b0 + j * 3 * 1024, 3 * 512, 128);
                } else {
                    ad[j] = make_desc(a0 + (j & 3) * 4160, 2080, 128);
                    bd[j] = make_desc(b0 + ((j + j / 4) & 3) * 3 * 1024, 3 * 512, 128);
                }
            }
        const long long t0 = clock64();
        uint32_t slot = 0;
        for (int r = 0; r < rows; ++r) {
            const uint32_t d0 = tmem_base + ((MODE & 4) ? slot * 32 : 0);
#pragma unroll
            for (int j = 0; j < 12; ++j) mma(d0, ad[j], bd[j], idesc);
            if (MODE & 32) commit(smem_u32(&scratch));
            if (++slot == 8) slot = 0;
            if (delay) {                      // per-row bookkeeping stand-in: does it overlap with the queued MMAs?
                const long long t = clock64();
                while (clock64() - t < delay) {}
            }
        }
        const long long t1 = clock64();
        commit(smem_u32(&bar));
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t2 = clock64();
        out[blockIdx.x] = t2 - t0;
        out[148 + blockIdx.x] = t1 - t0;
        stop = 1;
    } else if (warp == 4 && (MODE & 8)) {
        // producer look-alike: 8 bulk copies of 2 KB per "row" into a 4-stage ring beyond the operands
        if ((tid & 31) == 0) {
            uint32_t ph[4] = {0, 0, 0, 0};
            const uint32_t ring = smem_u32(smem) + 140 * 1024;
            int st = 0;
            size_t off = (size_t)blockIdx.x * (1 << 20);
            while (!stop) {
                const uint32_t b = smem_u32(&cbar[st]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(8u * 2048u) : "memory");
                for (int q = 0; q < 8; ++q)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(ring + st * 16384 + q * 2048), "l"(reinterpret_cast<const char*>(gsrc) + off + q * 2048), "r"(2048u), "r"(b) : "memory");
                off = (off + 16384) % ((size_t)120 << 20);
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(b), "r"(ph[st]) : "memory");
                ph[st] ^= 1u;
                st = (st + 1) & 3;
            }
        }
    } else if (warp >= 1 && warp <= 3 && (MODE & 16)) {
        // epilogue look-alike: read 32 columns, write 32 zeros, on TMEM columns the MMAs do not touch (lane quarter = warp)
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + 320;
        while (!stop) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                "tcgen05.wait::ld.sync.aligned;"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(ta) : "memory");
            const uint32_t z = r[0] & 0u;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
                "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n\t"
                "tcgen05.wait::st.sync.aligned;"
                ::"r"(ta), "r"(z) : "memory");
            __nanosleep(400);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// The conv kernel's issuer-loop skeleton around the same 12 MMAs, one element at a time (warp-uniform loop, elected lane
// issues).  SK bit 0: mbarrier try_wait on an already-complete phase + tcgen05.fence::after   bit 1: elect.sync +
// __syncwarp around the issue   bit 2: second elect + commit(empty)   bit 3: third elect + commit(acc_full)
// bit 4: the row bookkeeping arithmetic (tap range, slot ring position, wrap split)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int slot_of(int ro) { return 15 - (ro & 15); }
template <int SK>
__global__ void __launch_bounds__(160, 1) skeleton(int rows, long long* out, int nrows_chunk) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, ready, scratch[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 160) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&ready)), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&scratch[0])), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&scratch[1])), "r"(1u) : "memory");
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
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    if (warp == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 100 * 1024;
        const uint64_t a_desc0 = make_desc(0, RW * 16, 128), b_desc0 = make_desc(0, 3 * 512, 128);
        const long long t0 = clock64();
        int stage = 0;
        for (int r = 0; r < rows; ++r) {
            const int ri = r % (nrows_chunk + 2);
            int dy_lo = 0, ndy = 3, s0 = 0, n1 = 3, s1 = 0;
            if (SK & 16) {
                dy_lo = ri > nrows_chunk - 1 ? ri - (nrows_chunk - 1) : 0;
                const int dy_hi = min(2, ri);
                ndy = dy_hi - dy_lo + 1;
                const int ro_top = ri - dy_lo;
                s0 = slot_of(ro_top);
                n1 = min(ndy, 16 - s0);
            }
            if (SK & 1) {
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&ready)), "r"(1u) : "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (((SK & 2) ? elect_one() : (tid == 0)) && ndy > 0) {
                const uint32_t a_lo = (a0 + stage * 16640) >> 4;
                const uint32_t w_lo = (b0 >> 4) + dy_lo * 32;
                const uint32_t d0 = tmem_base + s0 * 32;
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((32 * n1) >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8)
                        mma(d0, a_desc0 | (uint64_t)(a_lo + dx + k8 * 2 * RW), b_desc0 | (uint64_t)(w_lo + (dx * 4 + k8) * 192), idesc);
                if (n1 < ndy) {
                    const uint32_t id1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((32 * (ndy - n1)) >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8)
                            mma(tmem_base + s1 * 32, a_desc0 | (uint64_t)(a_lo + dx + k8 * 2 * RW),
                                b_desc0 | (uint64_t)(w_lo + n1 * 32 + (dx * 4 + k8) * 192), id1);
                }
            }
            if (SK & 2) __syncwarp();
            if ((SK & 4) && elect_one()) commit(smem_u32(&scratch[0]));
            if (++stage == 4) stage = 0;
            if ((SK & 8) && elect_one()) commit(smem_u32(&scratch[1]));
        }
        if (tid == 0) {
            commit(smem_u32(&bar));
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            out[blockIdx.x] = clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
template <int SK>
void run_skeleton(long long* d_out, const char* what) {
    const int rows = 3000;
    cudaFuncSetAttribute(skeleton<SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int rep = 0; rep < 2; ++rep) skeleton<SK><<<148, 160, 210 * 1024>>>(rows, d_out, 69);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("skeleton %d: CUDA error %s\n", SK, cudaGetErrorString(e)); return; }
    std::vector<long long> t(148);
    cudaMemcpy(t.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
    std::sort(t.begin(), t.end());
    printf("skeleton %2d  %-66s %.0f cycles per row (12 MMAs = 672)\n", SK, what, (double)t[74] / rows);
}

template <int MODE>
void run(long long* d_out, const float* gsrc, const char* what) {
    const int rows = 3000;
    cudaFuncSetAttribute(ubench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int rep = 0; rep < 2; ++rep) ubench<MODE><<<148, 160, 210 * 1024>>>(rows, d_out, gsrc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", MODE, cudaGetErrorString(e)); return; }
    std::vector<long long> t(296);
    cudaMemcpy(t.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
    std::sort(t.begin(), t.begin() + 148);
    printf("mode %2d  %-70s %.1f cycles per MMA (N=96, 12 per row)\n", MODE, what, (double)t[74] / (rows * 12.0));
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 296 * 8);
    float* gsrc; cudaMalloc(&gsrc, (size_t)280 << 20); cudaMemset(gsrc, 0, (size_t)280 << 20);
    run<0>(d_out, gsrc, "ubench2-like descriptors");
    run<1>(d_out, gsrc, "kernel-like descriptors (16-byte tap shifts, one accumulator)");
    run<3>(d_out, gsrc, "kernel-like, 32-byte aligned tap shifts");
    run<5>(d_out, gsrc, "kernel-like, accumulator rotates per row");
    run<33>(d_out, gsrc, "kernel-like + commit per row");
    run<9>(d_out, gsrc, "kernel-like + concurrent bulk copies into shared memory");
    run<17>(d_out, gsrc, "kernel-like + concurrent tcgen05.ld/st");
    run<57>(d_out, gsrc, "kernel-like + commit + bulk copies + tcgen05.ld/st");
    // how much issuer-side work per row hides behind the MMA queue?
    cudaFuncSetAttribute(ubench<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int delay : {0, 50, 100, 150, 200, 300, 400, 600, 800}) {
        for (int rep = 0; rep < 2; ++rep) ubench<33><<<148, 160, 210 * 1024>>>(3000, d_out, gsrc, delay);
        cudaDeviceSynchronize();
        std::vector<long long> t(296);
        cudaMemcpy(t.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
        std::sort(t.begin(), t.begin() + 148);
        printf("issuer busy %3d cycles between rows of 12 MMAs (672 cycles of tensor work): %.0f cycles per row\n", delay, (double)t[74] / 3000.0);
    }
    run_skeleton<0>(d_out, "single thread, nothing but the MMAs");
    run_skeleton<2>(d_out, "warp-uniform loop, elect.sync + __syncwarp");
    run_skeleton<3>(d_out, "+ try_wait on a complete phase + tcgen05.fence::after");
    run_skeleton<7>(d_out, "+ elect + commit(empty)");
    run_skeleton<15>(d_out, "+ elect + commit(acc_full)");
    run_skeleton<31>(d_out, "+ row bookkeeping (tap range, slot, ring-wrap split)");
    return 0;
}
