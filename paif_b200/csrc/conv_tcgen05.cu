// tcgen05 TF32 implicit-GEMM convolution engine for sm_100a ("row-streaming" design).
//
// GEMM view per output row segment:  D[128 pixels x 32 cout] += A[128 pixels x 8 cin] * B[8 cin x 32 cout]
// for every (tap, 8-channel slice) — M=128, N=32, K=8 tcgen05.mma.kind::tf32, fp32 accumulators in TMEM.
//
//  * Activations are C4 maps ([B][C/4][H][W][4]): one quad-plane row IS a K-major, no-swizzle UMMA
//    operand (8 consecutive pixels x 16 B = one core matrix, SBO = 128 B, LBO = plane pitch), so a tap
//    (dy,dx) is just a 16*dx-byte shift of the descriptor start address: no im2col, no staging.
//  * One CTA owns a 128-pixel-wide column strip x RCH rows.  Producer warps stream halo'd input rows
//    (cp.async.bulk, one contiguous copy per quad plane) through a shared-memory ring; each input row is used by all
//    k*k taps on arrival: tap row dy accumulates into the TMEM accumulator of output row (ri - dy*dil + pad).
//    Accumulators form a 16-slot ring in TMEM (512 columns), so input is read once (+x/y halo).
//  * Weights (TF32-rounded, pre-packed as UMMA B tiles) stay resident in shared memory; when they do
//    not fit (7x7: 196 KB) K is split into passes over <=16 rows whose accumulators stay in TMEM.
//  * Warp roles: 0-7 epilogue in two groups that alternate output rows (TMEM -> registers -> fused
//    epilogue -> coalesced quad stores), 8-11 MMA issuers (one elected thread each; 8 also allocates TMEM), 12 producer
//    (one elected thread issuing cp.async.bulk row copies that complete on the stage's mbarrier).
//    mbarrier pipelines: full/empty per ring stage, acc_full/acc_empty per accumulator slot,
//    wfull/wempty for the weight slab.  Zero padding: rows outside the image are skipped (no copy, no
//    MMA), columns outside the image are zeroed once per stage and never overwritten.
#include "conv_epilogue.cuh"

namespace paif {

constexpr int TC_TW = 128;            // pixels per MMA (M)
constexpr int TC_SLOTS = 16;          // TMEM accumulator ring (16 x 32 columns = 512)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 8;                                   // two groups of 4 (TMEM lane quarters)
// One tcgen05.mma of this shape (M128 N32 K8) occupies its issuing thread for ~91-98 cycles although the
// tensor pipe needs 16 and the shared-memory operand fetch ~40 (scripts/mma_ubench.cu, measured on B200:
// 1 issuer 209 TFLOP/s, 2 issuers 389, 4 issuers 476 = operand-fetch bound).  So four warps issue
// concurrently; each owns the output rows (TMEM slots) with row % 4 == its index, which also keeps every
// accumulator's MMAs and its acc_full commit in one thread's program order.
constexpr int TC_MMA_WARPS = 4;
constexpr int TC_MMA_WARP = TC_EPI_WARPS, TC_PROD_WARP = TC_EPI_WARPS + TC_MMA_WARPS;
constexpr int TC_NT = (TC_EPI_WARPS + TC_MMA_WARPS + 1) * 32;     // 416
constexpr int TC_SMEM_BUDGET = 222 * 1024;
constexpr int TC_WSLAB_MAX = 110 * 1024;

struct TcPlan {
    int KQ;            // quads per pipeline unit (8 = whole source row, 4 = half)
    int gps;           // units (K-groups) per source = 8 / KQ
    int npass;         // K passes
    int gpp;           // K-groups per pass
    int RW;            // halo'd row width in pixels
    int pad;
    int unit_bytes;    // one ring stage
    int slab_bytes;    // resident weight slab per pass
    int stages;
    int smem_bytes;
};

static bool tc_make_plan(int nsrc, int k, int dil, TcPlan* p) {
    const int taps = k * k;
    p->pad = dil * (k - 1) / 2;
    p->RW = TC_TW + 2 * p->pad;
    const int full = nsrc * taps * 4096;                 // all weights (32 cin x 32 cout x 4 B per tap and source)
    if (full <= TC_WSLAB_MAX) { p->KQ = 8; p->npass = 1; p->gpp = nsrc; }
    else if (taps * 4096 <= TC_WSLAB_MAX) { p->KQ = 8; p->npass = nsrc; p->gpp = 1; }
    else if (taps * 2048 <= TC_WSLAB_MAX) { p->KQ = 4; p->npass = nsrc * 2; p->gpp = 1; }
    else return false;
    p->gps = 8 / p->KQ;
    p->slab_bytes = p->gpp * taps * p->KQ * 512;
    p->unit_bytes = p->KQ * p->RW * 16;
    int stages = (TC_SMEM_BUDGET - p->slab_bytes - 1024) / p->unit_bytes;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (stages < 2) return false;
    p->stages = stages;
    p->smem_bytes = p->slab_bytes + stages * p->unit_bytes + 1024;
    return true;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1ull << 22)) {          // a lost arrival would otherwise hang the GPU box
            printf("paif conv_tc: mbarrier wait timed out (block %d,%d,%d thread %d bar %u parity %u)\n",
                   blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// one elected lane of a converged warp (the rest of the warp-uniform control flow stays in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (async proxy)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::tf32, D=f32, A/B = TF32 K-major, N=32, M=128 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcGeom {
    int B, H, W, nsrc, k, dil, RCH, tiles_alloc;
    const float* src[3];
    const float* wmma;     // [K-group of KQ quads][tap][KQ/2 (k8)][2 (16-B chunk)][32 cout][4 cin], TF32-rounded
    TcPlan plan;
};

struct TcBars {
    uint64_t full[TC_MAX_STAGES];
    uint64_t empty[TC_MAX_STAGES];
    uint64_t acc_full[TC_SLOTS];
    uint64_t acc_empty[TC_SLOTS];
    uint64_t wfull, wempty;
    uint32_t tmem_base;
    uint32_t pad_;
};
static_assert(sizeof(TcBars) <= 1024, "barrier block must fit its 1 KB reservation");
static_assert(TC_SLOTS % TC_MMA_WARPS == 0, "slot ownership (row % TC_MMA_WARPS) must be stable across ring wraps");

__global__ void __launch_bounds__(TC_NT, 1)
conv_tc_kernel(TcGeom g, EpiParams e) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TcPlan& P = g.plan;
    unsigned char* s_w = smem;                                 // weight slab
    unsigned char* s_ring = smem + P.slab_bytes;               // input ring
    TcBars* bars = reinterpret_cast<TcBars*>(s_ring + P.stages * P.unit_bytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * TC_TW, r0 = blockIdx.y * g.RCH, b = blockIdx.z;
    const int nrows = min(g.RCH, g.H - r0);
    const int k = g.k, dil = g.dil, pad = P.pad, taps = k * k;
    const int nin = nrows + 2 * pad;                           // input rows touched per pass
    // valid x range of the halo'd row segment: columns [poff, poff + npx) of the RW-wide smem row
    const int xs = max(0, x0 - pad), xe = min(g.W, x0 + TC_TW + pad);
    const int poff = xs - (x0 - pad), npx = xe - xs;

    if (tid == 0) {
        for (int i = 0; i < TC_MAX_STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), TC_MMA_WARPS); }
        for (int i = 0; i < TC_SLOTS; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 128); }
        mbar_init(smem_u32(&bars->wfull), 1);
        mbar_init(smem_u32(&bars->wempty), TC_MMA_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (npx < P.RW) {
        // image-border strip: zero the columns no bulk copy will ever write (the conv's zero padding)
        const int nplanes = P.stages * P.KQ;
        const int nzero = P.RW - npx;
        for (int i = tid; i < nplanes * nzero; i += TC_NT) {
            const int pl = i / nzero, j = i - pl * nzero;
            const int px = j < poff ? j : j + npx;
            *reinterpret_cast<float4*>(s_ring + ((size_t)pl * P.RW + px) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);    // warp-uniform for the compiler

    if (warp < TC_EPI_WARPS) {
        // ===================== epilogue: TMEM -> registers -> fused epilogue -> global =====================
        const int grp = warp >> 2, wq = warp & 3;               // row-interleaved groups; TMEM lane quarter
        const int x = x0 + wq * 32 + lane;
        float csum[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) csum[c] = 0.f;
        for (int ro = grp; ro < nrows; ro += 2) {
            const int slot = ro % TC_SLOTS, use = ro / TC_SLOTS;
            mbar_wait(smem_u32(&bars->acc_full[slot]), use & 1);
            tc_fence_after();
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + slot * 32, v);
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->acc_empty[slot]));
            if (x < g.W) {
                epilogue_pixel<32>(e, b, r0 + ro, x, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) csum[c] += v[c];
            }
        }
        if (e.chan_partials) {
            // deterministic per-CTA channel sums: shuffle tree, then fixed-order cross-warp sum via smem
            float* red = reinterpret_cast<float*>(s_ring);          // the ring is idle once the last accumulator is done
            asm volatile("bar.sync 1, 256;" ::: "memory");            // both groups have drained their last rows
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float t = csum[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) red[warp * 32 + c] = t;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < 32) {
                float t = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < TC_EPI_WARPS; ++w8) t += red[w8 * 32 + tid];
                const int tile = blockIdx.y * gridDim.x + blockIdx.x;
                e.chan_partials[((size_t)b * g.tiles_alloc + tile) * 32 + tid] = t;
            }
        }
    } else if (warp < TC_PROD_WARP) {
        // ===================== MMA issuers (each warp runs the uniform loop; one elected lane issues) =====================
        {
            const int mw = warp - TC_MMA_WARP;                   // owns output rows with (row % TC_MMA_WARPS) == mw
            const uint32_t plane_bytes = P.RW * 16;
            const uint32_t w_base = smem_u32(s_w), ring_base = smem_u32(s_ring);
            const uint64_t a_desc0 = make_desc(0, plane_bytes, 128);
            const uint64_t b_desc0 = make_desc(0, 512, 128);
            const int nk8 = P.KQ / 2;
            uint32_t fresh = 0;                                  // slots whose next MMA must overwrite (not accumulate)
            int u = 0;                                           // ring position (valid units only)
            for (int pass = 0; pass < P.npass; ++pass) {
                mbar_wait(smem_u32(&bars->wfull), pass & 1);
                tc_fence_after();
                for (int ri = 0; ri < nin; ++ri) {               // input row y = r0 - pad + ri
                    const int y = r0 - pad + ri;
                    const bool yok = (y >= 0 && y < g.H);
                    for (int gl = 0; gl < P.gpp; ++gl) {
                        if (pass == 0 && gl == 0 && ri < nrows && (ri % TC_MMA_WARPS) == mw) {
                            // output row ri starts accumulating now: claim its TMEM slot
                            const int slot = ri % TC_SLOTS, use = ri / TC_SLOTS;
                            if (use > 0) {
                                mbar_wait(smem_u32(&bars->acc_empty[slot]), (use - 1) & 1);
                                tc_fence_after();
                            }
                            fresh |= 1u << slot;
                        }
                        if (yok) {
                            const int stage = u % P.stages;
                            mbar_wait(smem_u32(&bars->full[stage]), (u / P.stages) & 1);
                            tc_fence_after();
                            const uint32_t a_base = ring_base + stage * P.unit_bytes;
                            for (int dy = 0; dy < k; ++dy) {
                                const int ro = ri - dy * dil;    // output row (chunk-relative) fed by this tap row
                                if (ro < 0 || ro >= nrows || (ro % TC_MMA_WARPS) != mw) continue;
                                const int slot = ro % TC_SLOTS;
                                const uint32_t d_tmem = tmem_base + slot * 32;
                                uint32_t acc = (fresh >> slot) & 1u ? 0u : 1u;
                                fresh &= ~(1u << slot);
                                // descriptors differ only in the 14-bit start-address field (units of 16 B)
                                const uint32_t w_lo = (w_base + ((gl * taps + dy * k) * nk8) * 1024) >> 4;
                                const uint32_t a_lo = a_base >> 4;
                                if (elect_one()) {
                                    uint32_t wl = w_lo;
                                    for (int dx = 0; dx < k; ++dx) {
                                        uint32_t al = a_lo + dx * dil;
#pragma unroll 2
                                        for (int k8 = 0; k8 < nk8; ++k8) {
                                            tc_mma_tf32(d_tmem, a_desc0 | (uint64_t)al, b_desc0 | (uint64_t)wl, TC_IDESC, acc);
                                            acc = 1u;
                                            al += 2 * P.RW;              // two quad planes = 2 * RW * 16 B
                                            wl += 64;                    // 1 KB per B tile
                                        }
                                    }
                                }
                            }
                            if (elect_one()) tc_commit(smem_u32(&bars->empty[stage]));   // stage reusable once these MMAs retire
                            ++u;
                        }
                        if (pass == P.npass - 1 && gl == P.gpp - 1) {
                            const int rdone = ri - (k - 1) * dil;           // output row whose last tap row just passed
                            if (rdone >= 0 && rdone < nrows && (rdone % TC_MMA_WARPS) == mw && elect_one()) tc_commit(smem_u32(&bars->acc_full[rdone % TC_SLOTS]));
                        }
                    }
                }
                if (pass + 1 < P.npass && elect_one()) tc_commit(smem_u32(&bars->wempty));  // weights of this pass no longer read
            }
        }
        __syncwarp();
    } else {
        // ===================== producer: weight slab + halo'd input rows via cp.async.bulk =====================
        {
            const size_t plane = (size_t)g.H * g.W;
            const uint32_t row_bytes = (uint32_t)npx * 16;
            int u = 0;
            for (int pass = 0; pass < P.npass; ++pass) {
                if (pass > 0) mbar_wait(smem_u32(&bars->wempty), (pass - 1) & 1);
                {
                    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(g.wmma) + (size_t)pass * P.slab_bytes;
                    const uint32_t wdst = smem_u32(s_w), wbar = smem_u32(&bars->wfull);
                    if (elect_one()) {
                        mbar_expect_tx(wbar, (uint32_t)P.slab_bytes);
                        for (int off = 0; off < P.slab_bytes; off += 16384) {
                            const int n = min(16384, P.slab_bytes - off);
                            bulk_g2s(wdst + off, wsrc + off, (uint32_t)n, wbar);
                        }
                    }
                }
                for (int ri = 0; ri < nin; ++ri) {
                    const int y = r0 - pad + ri;
                    if (y < 0 || y >= g.H) continue;             // zero padding rows: skipped by the MMA issuer too
                    for (int gl = 0; gl < P.gpp; ++gl, ++u) {
                        const int stage = u % P.stages;
                        mbar_wait(smem_u32(&bars->empty[stage]), ((u / P.stages) & 1) ^ 1);
                        const int gk = pass * P.gpp + gl;        // global K-group
                        const int s = gk / P.gps, qoff = (gk % P.gps) * P.KQ;
                        const float4* sp = reinterpret_cast<const float4*>(g.src[s]) + ((size_t)b * 8 + qoff) * plane
                                           + (size_t)y * g.W + xs;
                        const uint32_t dst = smem_u32(s_ring) + stage * P.unit_bytes + poff * 16;
                        const uint32_t bar = smem_u32(&bars->full[stage]);
                        if (elect_one()) {
                            mbar_expect_tx(bar, row_bytes * P.KQ);
                            for (int q = 0; q < P.KQ; ++q) bulk_g2s(dst + q * P.RW * 16, sp + (size_t)q * plane, row_bytes, bar);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool conv_tc_supported(const PaifConvDesc& d) {
    if (d.cout != 32 || d.cin_per_src != 32 || d.kh != d.kw) return false;
    TcPlan p;
    return tc_make_plan(d.nsrc, d.kh, d.dil, &p);
}

static int tc_rows_per_cta(const PaifConvDesc& d, const TcPlan& p) {
    // multi-pass plans keep every output row of the chunk in TMEM (<= 16 slots)
    if (p.npass > 1) return TC_SLOTS;
    const int strips = cdiv(d.W, TC_TW);
    int rch = 32;
    // aim for >= 2 waves of 148 CTAs when the problem allows it
    while (rch > 8 && (long long)strips * cdiv(d.H, rch) * d.B < 2 * 148) rch >>= 1;
    return rch;
}

int conv_tc_tiles(int H, int W) {
    // upper bound independent of the per-launch row chunk: the smallest chunk is 8 rows
    return cdiv(W, TC_TW) * cdiv(H, 8);
}

int conv_tc_launch(const PaifConvDesc& d, cudaStream_t stream) {
    TcGeom g;
    if (!tc_make_plan(d.nsrc, d.kh, d.dil, &g.plan)) { set_error("conv_tc: no plan"); return PAIF_ENOTSUP; }
    g.B = d.B; g.H = d.H; g.W = d.W; g.nsrc = d.nsrc; g.k = d.kh; g.dil = d.dil;
    g.RCH = tc_rows_per_cta(d, g.plan);
    g.tiles_alloc = conv_tc_tiles(d.H, d.W);
    for (int i = 0; i < 3; ++i) g.src[i] = d.src[i];
    g.wmma = reinterpret_cast<const float*>(d.weight_mma);
    EpiParams e = make_epi(d);
    static int smem_set = 0;
    if (g.plan.smem_bytes > smem_set) {
        cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET + 1024);
        if (err != cudaSuccess) { set_error("conv_tc smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        smem_set = TC_SMEM_BUDGET + 1024;
    }
    dim3 grid(cdiv(d.W, TC_TW), cdiv(d.H, g.RCH), d.B);
    if (d.chan_partials) {
        // partial-sum slots beyond this launch's tile count must read as zero
        cudaError_t err = cudaMemsetAsync(d.chan_partials, 0, (size_t)d.B * conv_tc_tiles(d.H, d.W) * 32 * sizeof(float), stream);
        if (err != cudaSuccess) { set_error("conv_tc memset: %s", cudaGetErrorString(err)); return (int)err; }
    }
    conv_tc_kernel<<<grid, TC_NT, g.plan.smem_bytes, stream>>>(g, e);
    return check_launch("paif_conv_forward(tcgen05)");
}

int conv_tc_kq(int nsrc, int k, int dil) {
    TcPlan p;
    return tc_make_plan(nsrc, k, dil, &p) ? p.KQ : 0;
}

}  // namespace paif
