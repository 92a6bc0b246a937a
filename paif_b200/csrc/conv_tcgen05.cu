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
//    k*k taps on arrival.  Input row ri feeds output rows ri - dy*dil (dy = 0..k-1): their accumulators sit in
//    CONSECUTIVE 32-column TMEM slots, so all dy taps of one (dx, 8 input channels) are ONE tcgen05.mma with
//    N = 32*k (B tile rows = (dy, cout)).  tcgen05.mma issue is blocking — the issuing thread stays at most ~1 MMA
//    ahead of the pipe, an M128 K8 MMA costs max(32 + N/4, N/2) cycles, and every cycle of issuer-side work between
//    MMAs adds to the row time (scripts/mma_ubench2.cu, mma_ubench3.cu) — so few wide MMAs issued straight-line are
//    what keeps a single issuer from pacing the kernel.
//    Accumulators form a 16-slot ring in TMEM (512 columns), so input is read once (+x/y halo); slots are
//    zeroed by the epilogue after it drains them, every MMA accumulates.
//  * Weights (TF32-rounded, pre-packed as UMMA B tiles) stay resident in shared memory; when they do
//    not fit (7x7: 196 KB) K is split into passes over <=16 rows whose accumulators stay in TMEM.
//  * Warp roles: 0-7 epilogue in two groups that alternate output rows (TMEM -> registers -> fused
//    epilogue -> coalesced quad stores), 8 MMA issuer (one elected thread) + TMEM allocator, 9 producer
//    (one elected thread issuing cp.async.bulk row copies that complete on the stage's mbarrier).
//    mbarrier pipelines: full/empty per ring stage, acc_full/acc_empty per accumulator slot,
//    wfull/wempty for the weight slab.  Zero padding: rows outside the image are skipped (no copy, no
//    MMA), columns outside the image are zeroed once per stage and never overwritten.
#include "conv_epilogue.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace paif {

constexpr int TC_TW = 128;            // pixels per MMA (M)
constexpr int TC_SLOTS = 16;          // TMEM accumulator ring (16 x 32 columns = 512)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 8;                                   // two groups of 4 (TMEM lane quarters)
constexpr int TC_MMA_WARPS = 1;
constexpr int TC_MMA_WARP = TC_EPI_WARPS, TC_PROD_WARP = TC_EPI_WARPS + TC_MMA_WARPS;
constexpr int TC_NT = (TC_EPI_WARPS + TC_MMA_WARPS + 1) * 32;     // 320
constexpr int TC_SMEM_BUDGET = 227 * 1024;   // the sm_100 opt-in maximum of dynamic shared memory per CTA
constexpr int TC_WSLAB_MAX = 110 * 1024;
constexpr int TC_WSLAB_RESIDENT_MAX = 196 * 1024;   // fp32 7x7: both K halves resident next to a 3-stage half-row ring

struct TcPlan {
    int KQ;            // quads per pipeline unit (8 = whole source row, 4 = half)
    int gps;           // units (K-groups) per source = 8 / KQ
    int npass;         // K passes
    int gpp;           // K-groups per pass
    int RW;            // halo'd row width in pixels
    int pad;
    int unit_bytes;    // one K-group of one halo'd input row
    int upr;           // units per ring stage: gpp (a whole input row of the pass: one barrier round trip per row) or 1
    int stage_bytes;   // upr * unit_bytes
    int slab_bytes;    // resident weight slab per pass
    int stages;
    int smem_bytes;
};

// in_bf16: the sources are C8 bf16 maps (4 planes of 16-byte pixel vectors per 32-channel source instead of 8);
// the byte geometry of rows, stages and B tiles is the same, there are just half as many planes / K steps.
// cp: accumulator columns per output row (32 = the 32-cout convolutions; 16 = the 1-cout stem_out stencil, padded).
static bool tc_make_plan(int nsrc, int k, int dil, bool in_bf16, TcPlan* p, int cp = 32) {
    if (dil != 1 && dil != 2) return false;              // slot arithmetic uses shifts (log2 dil)
    if ((TC_SLOTS >> (dil >> 1)) < k) return false;      // the k rows fed by one input row need distinct slots
    const int taps = k * k;
    p->pad = dil * (k - 1) / 2;
    p->RW = TC_TW + 2 * p->pad;
    const int pps = in_bf16 ? 4 : 8;                     // planes per source
    const int src_bytes = taps * pps * 16 * cp;          // weights of one source (32 cin x cp cout per tap)
    if (nsrc * src_bytes <= TC_WSLAB_MAX) { p->KQ = pps; p->npass = 1; p->gpp = nsrc; }
    // fp32 7x7 32->32 (196 KB of weights): everything resident, half-row (4-plane) ring stages.  The layer is tensor-bound
    // (28 N=224 MMAs = 3136 cycles per row), so a ring of 1.5 rows covers the HBM latency, and a single pass means
    // tall chunks (no 16-row TMEM limit, 6 halo rows per chunk instead of per 16 rows) and no per-pass weight reload.
    else if (!in_bf16 && cp == 32 && k == 7 && nsrc == 1 && src_bytes <= TC_WSLAB_RESIDENT_MAX &&
             TC_SMEM_BUDGET - src_bytes - 1024 >= 3 * (pps / 2) * p->RW * 16) { p->KQ = pps / 2; p->npass = 1; p->gpp = 2; }
    else if (src_bytes <= TC_WSLAB_MAX) { p->KQ = pps; p->npass = nsrc; p->gpp = 1; }
    else if (!in_bf16 && src_bytes / 2 <= TC_WSLAB_MAX) { p->KQ = pps / 2; p->npass = nsrc * 2; p->gpp = 1; }
    else return false;
    p->gps = pps / p->KQ;
    p->slab_bytes = p->gpp * taps * p->KQ * 16 * cp;
    p->unit_bytes = p->KQ * p->RW * 16;
    // A ring stage holds a whole input row of the pass when at least 4 such rows fit (the MMA issuer then waits and
    // commits once per row instead of once per K-group); otherwise one K-group per stage, released as it retires.
    const int room = TC_SMEM_BUDGET - p->slab_bytes - 1024;
    // (1x1 layers are HBM-bound with 4 MMAs per K-group: finer release granularity wins there, measured 0.396 vs 0.415 ms)
    p->upr = (k > 1 && room / (p->gpp * p->unit_bytes) >= 4) ? p->gpp : 1;
    p->stage_bytes = p->upr * p->unit_bytes;
    int stages = room / p->stage_bytes;
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (stages < 2) return false;
    p->stages = stages;
    p->smem_bytes = p->slab_bytes + stages * p->stage_bytes + 1024;
    return true;
}

// TMEM slot of chunk-relative output row ro: rows of one dilation residue class descend through consecutive slots,
// so the k rows fed by one input row (ro, ro - dil, ...) occupy ascending consecutive slots.
// dsh = log2(dil), dil in {1, 2}: shifts and masks only (a runtime integer division costs ~100 cycles on the
// single thread that paces the tensor pipe).
#ifdef PAIF_TC_OLD_DESC      // A/B build: descriptors as (template | start address) with the plan fields read where they are used
#define desc_pack(lo, hi) (((uint64_t)(hi) << 32) | (uint64_t)(uint32_t)(lo))
#endif
__device__ __forceinline__ int tc_slot(int ro, int dsh) {
    const int spr = TC_SLOTS >> dsh;                 // slots per residue class ring
    return (ro & ((1 << dsh) - 1)) * spr + (spr - 1 - ((ro >> dsh) & (spr - 1)));
}

#ifdef PAIF_TC_PROFILE
// development-only role timeline: wait / busy cycles summed over all CTAs (read with paif_debug_tc_counters)
__device__ unsigned long long tc_prof[16];
#define TC_PROF_DECL long long prof_wait = 0, prof_wait2 = 0, prof_t0 = clock64()
#define TC_WAIT(stmt) do { const long long t_ = clock64(); stmt; prof_wait += clock64() - t_; } while (0)
#define TC_WAIT2(stmt) do { const long long t_ = clock64(); stmt; prof_wait2 += clock64() - t_; } while (0)   /* issuer: accumulator slot */
#define TC_PROF_END(slot) do { if ((threadIdx.x & 31) == 0) { atomicAdd(&tc_prof[(slot) * 2], (unsigned long long)prof_wait); \
                                   atomicAdd(&tc_prof[(slot) * 2 + 1], (unsigned long long)(clock64() - prof_t0));            \
                                   if (prof_wait2) atomicAdd(&tc_prof[8 + (slot)], (unsigned long long)prof_wait2); } } while (0)
#else
#define TC_PROF_DECL
#define TC_WAIT(stmt) stmt
#define TC_WAIT2(stmt) stmt
#define TC_PROF_END(slot)
#endif

struct TcGeom {
    int B, H, W, nsrc, k, dil, RCH, tiles_alloc;
    int epi_prefetch;      // epilogue threads L2-prefetch their residual / mask rows two rows ahead of the fetch
    int persist;           // 1: grid = (n, 1, 1) persistent CTAs, each owning an equal share of the launch's output rows
    int strips, rows_total;   // column strips per image; B * strips * H
    const void* src[3];
    const void* wmma;      // [K-group of KQ planes][dx][KQ/2 (K step)][2 (16-B chunk)][dy][32 cout][16 B of cin]:
                           // 4 TF32-rounded fp32 or 8 bf16 input channels per 16 bytes
    TcPlan plan;
};

struct TcBars {
    uint64_t full[TC_MAX_STAGES];
    uint64_t empty[TC_MAX_STAGES];
    uint64_t acc_full[TC_SLOTS];
    uint64_t acc_empty[TC_SLOTS];
    uint64_t wfull, wempty, zeroed;
    uint32_t tmem_base;
    alignas(16) float ch_scale[32];       // epilogue per-channel affine (1 / 0 when absent), read as float4
    alignas(16) float ch_shift[32];
};
static_assert(sizeof(TcBars) <= 1024, "barrier block must fit its 1 KB reservation");

struct TcSeg { int x0, r0, b, nrows, xs, poff, npx; };

// Block-wide barrier of the segment loops.  Each warp role runs its own (inlined) copy of tc_seg_begin, so the three roles
// arrive at three different bar.sync instructions.  The hardware barrier counts arriving warps whatever their program
// counter, but that is outside the documented contract of __syncthreads() and compute-sanitizer's synccheck reports it
// as "divergent thread(s) in block".  The sanitizer build (python -m paif_b200.build --sanitize: -DPAIF_SANITIZER_BUILD)
// therefore keeps the barrier in a non-inlined function — every thread executes the SAME instruction, synccheck is clean
// (profiles/r2_sanitizer_synccheck.log) — while the production build inlines it: the mere presence of a call in the
// kernel cost every conv layer 12-17 % (measured; values no longer stay in uniform registers across the issue loop).
#ifdef PAIF_SANITIZER_BUILD
__device__ __noinline__ void tc_block_sync() { __syncthreads(); }
#else
__device__ __forceinline__ void tc_block_sync() { __syncthreads(); }
#endif

// Start of a segment (see the kernel): geometry of the next piece of this CTA's share, and — block-wide — the drain of
// the previous segment, re-armed accumulator barriers, re-zeroed border columns.  Called by every thread of the CTA.
template <int K, int DIL, int KQ>
__device__ __forceinline__ void tc_seg_begin(const TcGeom& g, TcBars* bars, unsigned char* s_ring, int tid, int seg,
                                             int& lin, int lin_end, TcSeg& sg) {
    constexpr int pad = DIL * (K - 1) / 2, RW = TC_TW + 2 * pad;
    if (g.persist == 2) {                                      // banded: lin counts image rows, the strip is the CTA's own
        sg.b = lin / g.H;
        sg.r0 = lin - sg.b * g.H;
        sg.nrows = min(g.H - sg.r0, lin_end - lin);
        sg.x0 = (int)(blockIdx.x % (unsigned)g.strips) * TC_TW;
        lin += sg.nrows;
    } else if (g.persist) {
        const int si = lin / g.H;                              // (image, strip) index
        sg.r0 = lin - si * g.H;
        sg.nrows = min(g.H - sg.r0, lin_end - lin);
        sg.b = si / g.strips;
        sg.x0 = (si - sg.b * g.strips) * TC_TW;
        lin += sg.nrows;
    } else {
        sg.x0 = blockIdx.x * TC_TW; sg.r0 = blockIdx.y * g.RCH; sg.b = blockIdx.z;
        sg.nrows = min(g.RCH, g.H - sg.r0);
        lin = lin_end;
    }
    // valid x range of the halo'd row segment: columns [poff, poff + npx) of the RW-wide smem row
    sg.xs = max(0, sg.x0 - pad);
    const int xe = min(g.W, sg.x0 + TC_TW + pad);
    sg.poff = sg.xs - (sg.x0 - pad);
    sg.npx = xe - sg.xs;
    if (seg > 0) {
        // every role is done with the previous segment: all MMAs retired (the epilogue saw the last acc_full), no bulk
        // copy in flight (the issuer consumed every full stage), every drained accumulator slot is zero again
        tc_fence_before();
        tc_block_sync();
        if (tid == 0) {
            for (int i = 0; i < TC_SLOTS; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 128); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (sg.npx < RW) {
        // image-border strip: zero the columns no bulk copy will ever write (the conv's zero padding)
        const int nplanes = g.plan.stages * g.plan.upr * KQ;
        const int nzero = RW - sg.npx;
        for (int i = tid; i < nplanes * nzero; i += TC_NT) {
            const int pl = i / nzero, j = i - pl * nzero;
            const int px = j < sg.poff ? j : j + sg.npx;
            *reinterpret_cast<float4*>(s_ring + ((size_t)pl * RW + px) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    tc_block_sync();
    tc_fence_after();
}

// K, DIL, KQ are compile-time so that the MMA issue loop unrolls into straight-line code whose descriptors differ
// from a per-row base by immediates: the single issuing thread then sustains the tensor pipe's own rate
// (max(32 + N/4, N/2) cycles per MMA, scripts/mma_ubench2.cu) instead of ~150 cycles of address arithmetic per MMA.
// ST = storage mode: 0 = fp32 C4 maps in and out (TF32 MMAs); 1 = bf16 C8 maps in and out (bf16 MMAs);
// 2 = fp32 C4 sources (TF32 MMAs), bf16 C8 residuals / outputs (the layer that enters the bf16 part of the net).
// CP = accumulator columns per output row: 32 for the 32-cout convolutions; 16 = single-output mode (stem_out's merged
// 5x5 32->1 stencil with its cout padded to the smallest MMA N step): e.out / e.out_pre are [B][H][W] planes and the
// epilogue is PReLU + tanh of channel 0.
template <int K, int DIL, int KQ, bool PARTIALS, int ST, int CP = 32>
__global__ void __launch_bounds__(TC_NT, 1)
conv_tc_kernel(TcGeom g, EpiParams e) {
    static_assert(CP == 32 || (CP == 16 && !PARTIALS), "accumulator slot width");
    constexpr bool IN_BF = ST == 1, OUT_BF = ST != 0;
    constexpr int PPS_IN = IN_BF ? 4 : 8;                      // 16-byte planes per 32-channel source map
    extern __shared__ __align__(128) unsigned char smem[];
    const TcPlan& P = g.plan;
    unsigned char* s_w = smem;                                 // weight slab
    unsigned char* s_ring = smem + P.slab_bytes;               // input ring
    TcBars* bars = reinterpret_cast<TcBars*>(s_ring + P.stages * P.stage_bytes);  // (P.unit_bytes == UNIT, set by the host)

    const int tid = threadIdx.x, lane = tid & 31;
    // (warp-uniform for the compiler: the role branches and everything nested in them — the segment / pass / row loops
    //  and their barrier waits — then need no divergence bookkeeping; with a plain tid >> 5 the row loop of the MMA
    //  issuer carried BSSY / BSYNC pairs and BMOV spills of convergence barriers)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr int k = K, dil = DIL, pad = DIL * (K - 1) / 2;
    constexpr int dsh = DIL >> 1;                              // log2(dil), dil in {1, 2}
    constexpr int RW = TC_TW + 2 * pad;                        // halo'd row width in pixels
    constexpr int UNIT = KQ * RW * 16;                         // bytes of one ring stage

    if (tid == 0) {
        for (int i = 0; i < TC_MAX_STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), TC_MMA_WARPS); }
        for (int i = 0; i < TC_SLOTS; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), 128); }
        mbar_init(smem_u32(&bars->wfull), 1);
        mbar_init(smem_u32(&bars->wempty), TC_MMA_WARPS);
        mbar_init(smem_u32(&bars->zeroed), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid >= 32 && tid < 64) {
        bars->ch_scale[tid - 32] = e.ch_scale ? e.ch_scale[tid - 32] : 1.f;
        bars->ch_shift[tid - 32] = e.ch_shift ? e.ch_shift[tid - 32] : 0.f;
    }
    // Persistent launches (g.persist): the launch's B * strips * H output rows, linearised as (image, strip, row), are
    // split into gridDim.x equal contiguous shares — every SM gets the same number of rows whatever the batch, so there
    // is no partial last wave and one prologue per SM.  A share is walked as segments (the part of it inside one
    // image strip); between segments the CTA drains (block-wide sync), re-arms the accumulator barriers and re-zeroes
    // the border columns.  The ring barriers, the weights and the TMEM allocation live across segments.  Output bits do
    // not depend on where a share starts: every output row accumulates its taps in input-row order from a zeroed slot.
    // Banded form (g.persist == 2, the default): the B * H image rows are split into gridDim.x / strips equal bands and
    // the `strips` CTAs of a band walk the SAME rows, one column strip each, roughly in lockstep: neighbouring strips
    // then find the DRAM atoms that straddle their common border in L2.  With unrelated row ranges per CTA (the linear
    // form) ncu showed 8 % more DRAM reads than the tiled launch (1.395 vs 1.290 GB on the 64->32 layer).
    int lin = 0, lin_end = 1;
    if (g.persist == 2) {
        const int nb = gridDim.x / g.strips, band = blockIdx.x / g.strips;
        lin = (int)((long long)band * (g.B * g.H) / nb);
        lin_end = (int)((long long)(band + 1) * (g.B * g.H) / nb);
    } else if (g.persist) {
        lin = (int)((long long)blockIdx.x * g.rows_total / gridDim.x);
        lin_end = (int)((long long)(blockIdx.x + 1) * g.rows_total / gridDim.x);
    }
    // (each role runs its own segment loop, so that a role only carries the state it needs in registers; all three
    //  execute the same block-wide barriers in tc_seg_begin)
#define TC_SEG_LOOP_BEGIN                                                                                     \
    for (int seg = 0; lin < lin_end; ++seg) {                                                                 \
        TcSeg sg_;                                                                                            \
        tc_seg_begin<K, DIL, KQ>(g, bars, s_ring, tid, seg, lin, lin_end, sg_);                               \
        const int x0 = sg_.x0, r0 = sg_.r0, b = sg_.b, nrows = sg_.nrows, nin = sg_.nrows + 2 * pad;          \
        const int xs = sg_.xs, poff = sg_.poff, npx = sg_.npx;                                                \
        (void)x0; (void)r0; (void)b; (void)nrows; (void)nin; (void)xs; (void)poff; (void)npx;                              \
        if (seg == 0) tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);   /* warp-uniform for the compiler */
    uint32_t tmem_base = 0;

    if (warp < TC_EPI_WARPS) {
        // ===================== epilogue: TMEM -> registers -> fused epilogue -> global =====================
        float csum[PARTIALS ? 32 : 1];                          // (PARTIALS launches are tiled: one segment)
#pragma unroll
        for (int c = 0; c < (PARTIALS ? 32 : 1); ++c) csum[c] = 0.f;
        TC_SEG_LOOP_BEGIN
        const int grp = warp >> 2, wq = warp & 3;               // row-interleaved groups; TMEM lane quarter
        const int x = x0 + wq * 32 + lane;
        const bool xin = x < g.W;
        const size_t plane = (size_t)g.H * g.W;
        const float a = e.slope ? __ldg(e.slope) : 1.f;
        const float ma = e.mask_slope ? __ldg(e.mask_slope) : 0.f;
        const float a2 = e.slope2 ? __ldg(e.slope2) : 1.f;
        const bool any_post = e.post_res[0] || e.post_res[1] || e.post_res[2];
        const bool any_pre = e.pre_res[0] || e.pre_res[1];
        const bool any_fetch = any_post || any_pre || e.mask_src;
        // Residual and mask maps do not depend on the accumulator: for this thread's NEXT row they are fetched while that
        // row is still being accumulated, so global-load latency never sits between acc_full and the stores.
        // All loads of a fetch are issued BEFORE the first instruction that depends on any of them: the first
        // post-activation residual lands in pn, the first pre-activation residual — or, when the layer has none, the
        // SECOND post-activation residual — in qn, a third map in temporaries that are summed into pn afterwards.
        // (Summing map by map as the loads were issued serialised one DRAM latency per map: with three residual maps
        // the epilogue paced the kernel, the MMA issuer waiting 70-80 % of its time for an accumulator slot.)
        // bf16 maps (OUT_BF): the fetched rows stay PACKED (four 16-register slots: maps A, B, C and the rare D / E) and are
        // unpacked plane pair by plane pair where they are used — unpacked sums next to the accumulators did not fit in
        // the 168 registers a thread of this 10-warp CTA has.
        float4 pn[OUT_BF ? 1 : 8], qn[OUT_BF ? 1 : 8];
        uint4 sA[OUT_BF ? 4 : 1], sB[OUT_BF ? 4 : 1], sC[OUT_BF ? 4 : 1], sD[OUT_BF ? 4 : 1];
        uint32_t mbits = 0;
        const float* mapA = e.post_res[0];
        const float* mapB_ = any_pre ? e.pre_res[0] : e.post_res[1];
        const float* mapC_ = any_pre ? e.post_res[1] : e.post_res[2];       // summed into pn
        const float* mapD = any_pre ? e.post_res[2] : nullptr;             // (rare) summed into pn
        const float* mapE = any_pre ? e.pre_res[1] : nullptr;              // (rare) summed into qn
        // The epilogue's configuration as bits of ONE register, made opaque to the compiler: tested in the per-row code,
        // every `if (e.mask_src)` / `if (any_pre)` ... otherwise re-reads its kernel parameter from the constant bank
        // (LDCU -> ISETP -> BRA, ~30 such dependent triples per row inside the unrolled loops) — measured: the bf16
        // 3x3 32->32 layer 0.223 -> 0.149 ms without them.
        const float *mapB = mapB_, *mapC = mapC_;
        asm volatile("" : "+l"(mapB), "+l"(mapC));            // (selected once, not per use)
        uint32_t cfg = (mapA ? 1u : 0u) | (mapB ? 2u : 0u) | (mapC ? 4u : 0u) | (mapD ? 8u : 0u) | (mapE ? 16u : 0u) |
                       (any_pre ? 32u : 0u) | (e.mask_src ? 64u : 0u) | (e.slope ? 128u : 0u) | (e.out_pre ? 256u : 0u) |
                       (e.out_act2 ? 512u : 0u) | (any_fetch ? 1024u : 0u) |
                       ((!e.ch_scale && !e.ch_shift && !any_fetch && !e.out_pre && !e.out_act2 && e.post_scale == 1.f) ? 2048u : 0u) |
                       ((e.ch_scale || e.ch_shift) ? 4096u : 0u);
        asm volatile("" : "+r"(cfg));
        const bool fA = cfg & 1u, fB = cfg & 2u, fC = cfg & 4u, fD = cfg & 8u, fE = cfg & 16u, f_pre = cfg & 32u,
                   f_mask = cfg & 64u, f_slope = cfg & 128u, f_outpre = cfg & 256u, f_act2 = cfg & 512u, f_fetch = cfg & 1024u, f_plain = cfg & 2048u, f_aff = cfg & 4096u;
        // L2 prefetch (per thread, no registers held) of the rows this thread will FETCH next time (two rows further down):
        // when the epilogue is the pacing role — the bf16 layers — its fetch is consumed almost immediately, i.e. with the
        // full DRAM latency exposed; with the lines already in L2 the exposed latency is an L2 hit.  bf16 maps only
        // (measured, same box: bf16 residual layers 7-10 % faster; fp32 layers, which run at 6 TB/s, -3 ... +9 %).
        const bool tc_epi_prefetch = OUT_BF && g.epi_prefetch != 0;
        auto prefetch_rows = [&](int ro) {
            constexpr int NP = OUT_BF ? 4 : 8;
            const size_t base = (size_t)b * NP * plane + (size_t)(r0 + ro) * g.W + x;        // 16-byte units, plane 0
            auto pf = [&](const float* m) {
#pragma unroll
                for (int pl = 0; pl < NP; ++pl)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(m) + base + pl * plane));
            };
            if (fA) pf(mapA);
            if (fB) pf(mapB);
            if (fC) pf(mapC);
            if (f_mask) pf(e.mask_src);
        };
        auto fetch_next = [&](int ro) {
            if (tc_epi_prefetch && ro + 2 < nrows) prefetch_rows(ro + 2);
            if constexpr (OUT_BF) {
                const size_t base = (size_t)b * 4 * plane + (size_t)(r0 + ro) * g.W + x;     // 16-byte units, plane 0
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                const float* mapX = fD ? mapD : mapE;                  // (the launcher rejects both at once)
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) sA[pl] = fA ? ld_stream_u4(reinterpret_cast<const uint4*>(mapA) + base + pl * plane) : z4;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) sB[pl] = fB ? ld_stream_u4(reinterpret_cast<const uint4*>(mapB) + base + pl * plane) : z4;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) sC[pl] = fC ? ld_stream_u4(reinterpret_cast<const uint4*>(mapC) + base + pl * plane) : z4;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) sD[pl] = (fD || fE) ? ld_stream_u4(reinterpret_cast<const uint4*>(mapX) + base + pl * plane) : z4;
                if (f_mask) {
                    uint4 rm[4];
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) rm[pl] = ld_stream_u4(reinterpret_cast<const uint4*>(e.mask_src) + base + pl * plane);
                    uint32_t mb = 0;
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl)
                        mb |= (bf2_pos_bits(rm[pl].x) | bf2_pos_bits(rm[pl].y) << 2 | bf2_pos_bits(rm[pl].z) << 4 | bf2_pos_bits(rm[pl].w) << 6) << (8 * pl);
                    mbits = mb;
                }
            } else {
            const size_t base = (size_t)b * 8 * plane + (size_t)(r0 + ro) * g.W + x;
            float4 tc[8];
            if (fA) {
#pragma unroll
                for (int q = 0; q < 8; ++q) pn[q] = ld_stream(reinterpret_cast<const float4*>(mapA) + base + q * plane);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) pn[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (fB) {
#pragma unroll
                for (int q = 0; q < 8; ++q) qn[q] = ld_stream(reinterpret_cast<const float4*>(mapB) + base + q * plane);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) qn[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (fC) {
#pragma unroll
                for (int q = 0; q < 8; ++q) tc[q] = ld_stream(reinterpret_cast<const float4*>(mapC) + base + q * plane);
            }
            if (fC) {                   // (a third map and a mask source do not occur together: they share registers)
#pragma unroll
                for (int q = 0; q < 8; ++q) pn[q] = f4_add(pn[q], tc[q]);
            }
            if (f_mask) {
                uint32_t mb = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 m = ld_stream(reinterpret_cast<const float4*>(e.mask_src) + base + q * plane);
                    mb |= (m.x > 0.f ? 1u : 0u) << (4 * q) | (m.y > 0.f ? 2u : 0u) << (4 * q) |
                          (m.z > 0.f ? 4u : 0u) << (4 * q) | (m.w > 0.f ? 8u : 0u) << (4 * q);
                }
                mbits = mb;
            }
            if (fD) {
#pragma unroll
                for (int q = 0; q < 8; ++q) pn[q] = f4_add(pn[q], ld_stream(reinterpret_cast<const float4*>(mapD) + base + q * plane));
            }
            if (fE) {
#pragma unroll
                for (int q = 0; q < 8; ++q) qn[q] = f4_add(qn[q], ld_stream(reinterpret_cast<const float4*>(mapE) + base + q * plane));
            }
            }
        };
#pragma unroll
        for (int q = 0; q < (OUT_BF ? 1 : 8); ++q) { pn[q] = make_float4(0.f, 0.f, 0.f, 0.f); qn[q] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int pl = 0; pl < (OUT_BF ? 4 : 1); ++pl) sA[pl] = sB[pl] = sC[pl] = sD[pl] = make_uint4(0u, 0u, 0u, 0u);
        if (f_fetch && xin && grp < nrows) fetch_next(grp);
        if (grp == 0 && seg == 0) {
            // accumulators start from zero: every MMA accumulates (the TMEM allocation holds garbage)
            for (int sl = 0; sl < TC_SLOTS; ++sl) {
                if constexpr (CP == 16) tmem_zero16(tmem_base + ((uint32_t)(wq * 32) << 16) + sl * CP);
                else tmem_zero32(tmem_base + ((uint32_t)(wq * 32) << 16) + sl * 32);
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->zeroed));
        }
        TC_PROF_DECL;
        for (int ro = grp; ro < nrows; ro += 2) {
            const int slot = tc_slot(ro, dsh), use = ro / TC_SLOTS;
            TC_WAIT(mbar_wait(smem_u32(&bars->acc_full[slot]), use & 1));
            tc_fence_after();
            if constexpr (CP == 16) {
                const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + slot * CP;
                const float acc = tmem_ld16_first(ta);
                tmem_zero16(ta);
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->acc_empty[slot]));
                if (xin) {
                    const size_t pi = (size_t)b * plane + (size_t)(r0 + ro) * g.W + x;
                    if (f_outpre) e.out_pre[pi] = acc;
                    e.out[pi] = tanhf(prelu_f(acc, a));
                }
                continue;
            }
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + slot * 32, v);
            tmem_zero32(tmem_base + ((uint32_t)(wq * 32) << 16) + slot * 32);
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->acc_empty[slot]));
            if (f_plain) {
                // out = PReLU(acc) (or acc): no affine, no residual, no second output — conv1 / conv2 of every dense
                // block.  A third of the general path's instructions; the epilogue warps are what paces the bf16 layers.
                if (xin) {
                    if (f_slope) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = prelu_f(v[c], a);
                    }
                    if constexpr (OUT_BF) {
                        uint4* op = reinterpret_cast<uint4*>(e.out) + (size_t)b * 4 * plane + (size_t)(r0 + ro) * g.W + x;
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl)
                            op[pl * plane] = bf8_pack(make_float4(v[pl * 8 + 0], v[pl * 8 + 1], v[pl * 8 + 2], v[pl * 8 + 3]),
                                                      make_float4(v[pl * 8 + 4], v[pl * 8 + 5], v[pl * 8 + 6], v[pl * 8 + 7]));
                    } else {
                        float4* op = reinterpret_cast<float4*>(e.out) + (size_t)b * 8 * plane + (size_t)(r0 + ro) * g.W + x;
#pragma unroll
                        for (int q = 0; q < 8; ++q) op[q * plane] = make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    }
                    if (PARTIALS) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) csum[c] += v[c];
                    }
                }
                continue;
            }
            if constexpr (OUT_BF) if (xin) {
                const size_t base = (size_t)b * 4 * plane + (size_t)(r0 + ro) * g.W + x;     // 16-byte units, plane 0
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    float tp[8];
                    // post- (P) and pre-activation (Q) residual sums of this plane's 8 channels from the packed slots
                    float4 P[2], Q[2];
                    P[0] = P[1] = Q[0] = Q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (fA) bf8_unpack(sA[pl], P[0], P[1]);
                    if (f_pre) bf8_unpack(sB[pl], Q[0], Q[1]);
                    else if (fB) { float4 lo, hi; bf8_unpack(sB[pl], lo, hi); P[0] = f4_add(P[0], lo); P[1] = f4_add(P[1], hi); }
                    if (fC) { float4 lo, hi; bf8_unpack(sC[pl], lo, hi); P[0] = f4_add(P[0], lo); P[1] = f4_add(P[1], hi); }
                    if (fD) { float4 lo, hi; bf8_unpack(sD[pl], lo, hi); P[0] = f4_add(P[0], lo); P[1] = f4_add(P[1], hi); }
                    else if (fE) { float4 lo, hi; bf8_unpack(sD[pl], lo, hi); Q[0] = f4_add(Q[0], lo); Q[1] = f4_add(Q[1], hi); }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * pl + h;
                        float t[4] = {v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]};
                        if (f_aff) {
                            const float4 sc = *reinterpret_cast<const float4*>(&bars->ch_scale[q * 4]);
                            const float4 sh = *reinterpret_cast<const float4*>(&bars->ch_shift[q * 4]);
                            t[0] = fmaf(t[0], sc.x, sh.x); t[1] = fmaf(t[1], sc.y, sh.y);
                            t[2] = fmaf(t[2], sc.z, sh.z); t[3] = fmaf(t[3], sc.w, sh.w);
                        }
                        if (f_pre) { t[0] += Q[h].x; t[1] += Q[h].y; t[2] += Q[h].z; t[3] += Q[h].w; }
#pragma unroll
                        for (int j = 0; j < 4; ++j) tp[h * 4 + j] = t[j];
                        if (f_mask) {
                            const uint32_t mq = mbits >> (4 * q);
                            t[0] *= (mq & 1u) ? 1.f : ma; t[1] *= (mq & 2u) ? 1.f : ma;
                            t[2] *= (mq & 4u) ? 1.f : ma; t[3] *= (mq & 8u) ? 1.f : ma;
                        } else if (f_slope) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) t[j] = prelu_f(t[j], a);
                        }
                        v[q * 4 + 0] = fmaf(t[0], e.post_scale, P[h].x); v[q * 4 + 1] = fmaf(t[1], e.post_scale, P[h].y);
                        v[q * 4 + 2] = fmaf(t[2], e.post_scale, P[h].z); v[q * 4 + 3] = fmaf(t[3], e.post_scale, P[h].w);
                    }
                    if (f_outpre)
                        reinterpret_cast<uint4*>(e.out_pre)[base + pl * plane] =
                            bf8_pack(make_float4(tp[0], tp[1], tp[2], tp[3]), make_float4(tp[4], tp[5], tp[6], tp[7]));
                }
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const float* w8 = &v[pl * 8];
                    reinterpret_cast<uint4*>(e.out)[base + pl * plane] =
                        bf8_pack(make_float4(w8[0], w8[1], w8[2], w8[3]), make_float4(w8[4], w8[5], w8[6], w8[7]));
                    if (f_act2)
                        reinterpret_cast<uint4*>(e.out_act2)[base + pl * plane] =
                            bf8_pack(make_float4(prelu_f(w8[0], a2), prelu_f(w8[1], a2), prelu_f(w8[2], a2), prelu_f(w8[3], a2)),
                                     make_float4(prelu_f(w8[4], a2), prelu_f(w8[5], a2), prelu_f(w8[6], a2), prelu_f(w8[7], a2)));
                }
                if (PARTIALS) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) csum[c] += v[c];
                }
                if (f_fetch && ro + 2 < nrows) fetch_next(ro + 2);      // in flight during the next row's MMAs
            }
            if constexpr (!OUT_BF) if (xin) {
                const size_t base = (size_t)b * 8 * plane + (size_t)(r0 + ro) * g.W + x;     // float4 units, quad 0
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const size_t off = base + q * plane;
                    float t[4] = {v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]};
                    if (f_aff) {
                        const float4 sc = *reinterpret_cast<const float4*>(&bars->ch_scale[q * 4]);
                        const float4 sh = *reinterpret_cast<const float4*>(&bars->ch_shift[q * 4]);
                        t[0] = fmaf(t[0], sc.x, sh.x); t[1] = fmaf(t[1], sc.y, sh.y);
                        t[2] = fmaf(t[2], sc.z, sh.z); t[3] = fmaf(t[3], sc.w, sh.w);
                    }
                    if (f_pre) { t[0] += qn[q].x; t[1] += qn[q].y; t[2] += qn[q].z; t[3] += qn[q].w; }
                    if (f_outpre) reinterpret_cast<float4*>(e.out_pre)[off] = make_float4(t[0], t[1], t[2], t[3]);
                    if (f_mask) {
                        const uint32_t mq = mbits >> (4 * q);
                        t[0] *= (mq & 1u) ? 1.f : ma; t[1] *= (mq & 2u) ? 1.f : ma;
                        t[2] *= (mq & 4u) ? 1.f : ma; t[3] *= (mq & 8u) ? 1.f : ma;
                    } else if (f_slope) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = prelu_f(t[j], a);
                    }
                    v[q * 4 + 0] = fmaf(t[0], e.post_scale, pn[q].x); v[q * 4 + 1] = fmaf(t[1], e.post_scale, pn[q].y);
                    v[q * 4 + 2] = fmaf(t[2], e.post_scale, pn[q].z); v[q * 4 + 3] = fmaf(t[3], e.post_scale, pn[q].w);
                    if (!f_pre) { v[q * 4 + 0] += qn[q].x; v[q * 4 + 1] += qn[q].y; v[q * 4 + 2] += qn[q].z; v[q * 4 + 3] += qn[q].w; }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const size_t off = base + q * plane;
                    reinterpret_cast<float4*>(e.out)[off] = make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    if (f_act2)
                        reinterpret_cast<float4*>(e.out_act2)[off] =
                            make_float4(prelu_f(v[q * 4 + 0], a2), prelu_f(v[q * 4 + 1], a2), prelu_f(v[q * 4 + 2], a2), prelu_f(v[q * 4 + 3], a2));
                }
                if (PARTIALS) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) csum[c] += v[c];
                }
                if (f_fetch && ro + 2 < nrows) fetch_next(ro + 2);      // in flight during the next row's MMAs
            }
        }
        TC_PROF_END(0);
        }   // segments
        if (PARTIALS) {
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
                e.chan_partials[((size_t)blockIdx.z * g.tiles_alloc + tile) * 32 + tid] = t;
            }
        }
    } else if (warp < TC_PROD_WARP) {
        // ===================== MMA issuer (the warp runs the uniform loop; one elected lane issues) =====================
        int stage = 0;                                           // ring position (valid units only) and its phase,
        uint32_t phase = 0;                                      // carried across segments
        TC_SEG_LOOP_BEGIN
        {
            constexpr uint32_t plane_bytes = RW * 16;
            const uint32_t w_base = smem_u32(s_w), ring_base = smem_u32(s_ring);
            const uint64_t a_desc0 = make_desc(0, plane_bytes, 128);
            const uint64_t b_desc0 = make_desc(0, (uint32_t)k * CP * 16, 128);   // 16-B k-chunks are k*CP rows apart
            // descriptor words: the start-address field (14 bits, 16-byte units) never carries into the LBO field, so a
            // descriptor is (high word, low-word base + row offset + compile-time tap offset): one add per MMA operand
            const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
            constexpr int nk8 = KQ / 2;
            constexpr int spr = TC_SLOTS >> dsh;
            // plan fields used per row: pinned in registers (the compiler otherwise re-reads them from the constant bank
            // on the issuing thread's critical path)
            int p_upr = P.upr, p_stages = P.stages, p_gpp = P.gpp, p_npass = P.npass;
            uint32_t p_stage16 = (uint32_t)P.stage_bytes >> 4;
#ifndef PAIF_TC_OLD_DESC
            asm volatile("" : "+r"(p_upr), "+r"(p_stages), "+r"(p_gpp), "+r"(p_npass), "+r"(p_stage16));
#endif
            const uint32_t ring16 = ring_base >> 4, w16 = w_base >> 4;
            TC_PROF_DECL;
            if (seg == 0) {
                TC_WAIT(mbar_wait(smem_u32(&bars->zeroed), 0));
                tc_fence_after();
            }
            for (int pass = 0; pass < p_npass; ++pass) {
                if (seg == 0) {                                  // (persistent launches are single-pass: the slab stays)
                    TC_WAIT(mbar_wait(smem_u32(&bars->wfull), pass & 1));
                    tc_fence_after();
                }
                for (int ri = 0; ri < nin; ++ri) {               // input row y = r0 - pad + ri
                    const int y = r0 - pad + ri;
                    const bool yok = (y >= 0 && y < g.H);
                    // taps dy whose output row ro = ri - dy*dil lies in the chunk: dy_lo..dy_hi, slots ascending from s0
                    const int dy_lo = ri > nrows - 1 ? (ri - (nrows - 1) + dil - 1) >> dsh : 0;
                    const int dy_hi = min(k - 1, ri >> dsh);
                    const int ndy = dy_hi - dy_lo + 1;
                    int s0 = 0, n1 = 0, s1 = 0;
                    if (ndy > 0) {
                        const int ro_top = ri - dy_lo * dil;
                        s0 = tc_slot(ro_top, dsh);
                        s1 = (ro_top & (dil - 1)) * spr;                       // start of this residue class's slot ring
                        n1 = min(ndy, s1 + spr - s0);                          // taps before the ring wraps
                    }
                    for (int gl = 0; gl < p_gpp; ++gl) {
                        if (pass == 0 && gl == 0 && ri < nrows) {
                            // output row ri starts accumulating now: its TMEM slot must have been drained and zeroed
                            const int use = ri / TC_SLOTS;
                            if (use > 0) {
                                TC_WAIT2(mbar_wait(smem_u32(&bars->acc_empty[tc_slot(ri, dsh)]), (use - 1) & 1));
                                tc_fence_after();
                            }
                        }
                        if (yok) {
                            const int us = p_upr == 1 ? 0 : gl;                 // unit inside the ring stage
                            if (us == 0) {
                                TC_WAIT(mbar_wait(smem_u32(&bars->full[stage]), phase));
                                tc_fence_after();
                            }
                            if (ndy > 0 && elect_one()) {
                                // descriptors differ only in the 14-bit start-address field (units of 16 B); the
                                // per-(dx, k8) offsets below are immediates after unrolling
                                const uint32_t a_lo = a_lo0 + ring16 + (uint32_t)stage * p_stage16 + (uint32_t)us * (UNIT >> 4);
                                const uint32_t w_lo = b_lo0 + w16 + (uint32_t)gl * (k * nk8 * k * CP * 2) + dy_lo * CP;
                                const uint32_t d0 = tmem_base + s0 * CP;
                                const uint32_t id0 = tc_idesc((uint32_t)CP * n1, IN_BF ? 1u : 2u);
#pragma unroll
                                for (int dx = 0; dx < k; ++dx)
#pragma unroll
                                    for (int k8 = 0; k8 < nk8; ++k8) {
                                        const uint64_t ad = desc_pack(a_lo + (dx * dil + k8 * 2 * RW), a_hi);
                                        const uint64_t bd = desc_pack(w_lo + ((dx * nk8 + k8) * k * 2 * CP), b_hi);
                                        if constexpr (IN_BF) tc_mma_bf16(d0, ad, bd, id0, 1u);
                                        else tc_mma_tf32(d0, ad, bd, id0, 1u);
                                    }
                                if (n1 < ndy) {                           // the slot ring wrapped: remaining taps start at s1
                                    const uint32_t d1 = tmem_base + s1 * CP;
                                    const uint32_t id1 = tc_idesc((uint32_t)CP * (ndy - n1), IN_BF ? 1u : 2u);
                                    const uint32_t w_l1 = w_lo + n1 * CP;
#pragma unroll
                                    for (int dx = 0; dx < k; ++dx)
#pragma unroll
                                        for (int k8 = 0; k8 < nk8; ++k8) {
                                            const uint64_t ad = desc_pack(a_lo + (dx * dil + k8 * 2 * RW), a_hi);
                                            const uint64_t bd = desc_pack(w_l1 + ((dx * nk8 + k8) * k * 2 * CP), b_hi);
                                            if constexpr (IN_BF) tc_mma_bf16(d1, ad, bd, id1, 1u);
                                            else tc_mma_tf32(d1, ad, bd, id1, 1u);
                                        }
                                }
                            }
                            __syncwarp();
                            if (us == p_upr - 1) {
                                if (elect_one()) tc_commit(smem_u32(&bars->empty[stage]));   // stage reusable once these MMAs retire
                                if (++stage == p_stages) { stage = 0; phase ^= 1u; }
                            }
                        }
                        if (pass == p_npass - 1 && gl == p_gpp - 1) {
                            const int rdone = ri - (k - 1) * dil;           // output row whose last tap row just passed
                            if (rdone >= 0 && rdone < nrows && elect_one()) tc_commit(smem_u32(&bars->acc_full[tc_slot(rdone, dsh)]));
                        }
                    }
                }
                if (pass + 1 < p_npass && elect_one()) tc_commit(smem_u32(&bars->wempty));  // weights of this pass no longer read
            }
            TC_PROF_END(1);
        }
        }   // segments
        __syncwarp();
    } else {
        // ===================== producer: weight slab + halo'd input rows via cp.async.bulk =====================
        int stage = 0;
        uint32_t phase = 0;
        TC_SEG_LOOP_BEGIN
        {
            const size_t plane = (size_t)g.H * g.W;
            const uint32_t row_bytes = (uint32_t)npx * 16;
            TC_PROF_DECL;
            for (int pass = 0; pass < P.npass; ++pass) {
                if (pass > 0) mbar_wait(smem_u32(&bars->wempty), (pass - 1) & 1);
                if (seg == 0) {
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
                    for (int gl = 0; gl < P.gpp; ++gl) {
                        const int us = P.upr == 1 ? 0 : gl;      // unit inside the ring stage
                        if (us == 0) {
                            TC_WAIT(mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1u));
                            if (elect_one()) mbar_expect_tx(smem_u32(&bars->full[stage]), row_bytes * KQ * P.upr);
                        }
                        const int gk = pass * P.gpp + gl;        // global K-group
                        constexpr int GPS = PPS_IN / KQ;         // K-groups per source (1 or 2)
                        const int s = GPS == 1 ? gk : gk >> 1, qoff = GPS == 1 ? 0 : (gk & 1) * KQ;
                        const float4* sp = reinterpret_cast<const float4*>(g.src[s]) + ((size_t)b * PPS_IN + qoff) * plane
                                           + (size_t)y * g.W + xs;
                        const uint32_t dst = smem_u32(s_ring) + stage * P.stage_bytes + us * UNIT + poff * 16;
                        const uint32_t bar = smem_u32(&bars->full[stage]);
                        if (elect_one()) {
#pragma unroll
                            for (int q = 0; q < KQ; ++q) bulk_g2s(dst + q * RW * 16, sp + (size_t)q * plane, row_bytes, bar);
                        }
                        if (us == P.upr - 1 && ++stage == P.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            TC_PROF_END(2);
        }
        }   // segments
        __syncwarp();
    }
#undef TC_SEG_LOOP_BEGIN

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Fused DilConv on the tensor pipe (operations_m.py:494-506):
//   out = ch_scale * pw1x1(dw3x3_dil(relu(x))) + ch_shift + (add_x ? x : 0) + r1 + r2
// One CTA = 128 pixels of a row strip x DC_ROWS rows, 4 warps, thread = pixel.  Per row: every thread computes
// the 32 depthwise results of its pixel (FFMA, 9 taps) and writes them to shared memory in the quad-plane
// layout, which IS the UMMA A operand; one thread issues 4 tcgen05.mma (M128 N32 K8, TF32) against the
// resident BN-scaled 1x1 weights; the same threads read their pixel's 32 outputs back from TMEM, add the
// shift and the residual maps and store.  Rows are serial inside a CTA; ~5 CTAs per SM overlap each other.
// ---------------------------------------------------------------------------------------------
constexpr int DC_ROWS = 16;

template <int DIL>
__global__ void __launch_bounds__(128)
dilconv_tc_kernel(const float* __restrict__ xin, const float* __restrict__ dw, const float* __restrict__ pw,
                  const float* __restrict__ ch_scale, const float* __restrict__ ch_shift,
                  const float* __restrict__ r1, const float* __restrict__ r2, float* __restrict__ out,
                  int add_x, int H, int W) {
    __shared__ __align__(128) float4 s_a[8 * 128];          // A operand: [quad plane][pixel][4 ch]
    __shared__ __align__(128) float s_b[4 * 2 * 32 * 4];    // B operand: [k8][16-B chunk][cout][4 cin], TF32, BN-scaled
    __shared__ __align__(16) float s_dw[9 * 32];            // [tap][channel]
    __shared__ __align__(16) float s_sh[32];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int x = blockIdx.x * 128 + tid, r0 = blockIdx.y * DC_ROWS, b = blockIdx.z;
    const int nrows = min(DC_ROWS, H - r0);
    for (int i = tid; i < 1024; i += 128) {
        const int co = i >> 5, ci = i & 31;
        const float w = pw[co * 32 + ci] * (ch_scale ? ch_scale[co] : 1.f);
        const uint32_t bits = (__float_as_uint(w) + 0x1000u) & 0xffffe000u;        // round to nearest TF32
        s_b[(((ci >> 3) * 2 + ((ci >> 2) & 1)) * 32 + co) * 4 + (ci & 3)] = __uint_as_float(bits);
    }
    for (int i = tid; i < 288; i += 128) { const int t = i >> 5, c = i & 31; s_dw[i] = dw[c * 9 + t]; }
    if (tid < 32) s_sh[tid] = ch_shift ? ch_shift[tid] : 0.f;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const size_t plane = (size_t)H * W;
    const float4* xp = reinterpret_cast<const float4*>(xin) + (size_t)b * 8 * plane;
    const bool xin_ok = x < W;
    const uint64_t a_desc0 = make_desc(smem_u32(s_a), 128 * 16, 128);
    const uint64_t b_desc0 = make_desc(smem_u32(s_b), 512, 128);
    uint32_t phase = 0;
    for (int ro = 0; ro < nrows; ++ro) {
        const int y = r0 + ro;
        // ---- depthwise 3x3 (dilation DIL) of relu(x): this thread's pixel, all 32 channels -> A operand
        int toff[9];
        float tval[9];
#pragma unroll
        for (int ty = 0; ty < 3; ++ty)
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
                const int yy = y + (ty - 1) * DIL, xx = x + (tx - 1) * DIL;
                const bool ok = xin_ok && yy >= 0 && yy < H && xx >= 0 && xx < W;
                toff[ty * 3 + tx] = ok ? yy * W + xx : 0;
                tval[ty * 3 + tx] = ok ? 1.f : 0.f;
            }
#pragma unroll 2
        for (int q = 0; q < 8; ++q) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 v = __ldg(xp + (size_t)q * plane + toff[t]);
                const float4 ww = *reinterpret_cast<const float4*>(&s_dw[t * 32 + q * 4]);
                const float m = tval[t];
                t4.x = fmaf(fmaxf(v.x, 0.f) * m, ww.x, t4.x); t4.y = fmaf(fmaxf(v.y, 0.f) * m, ww.y, t4.y);
                t4.z = fmaf(fmaxf(v.z, 0.f) * m, ww.z, t4.z); t4.w = fmaf(fmaxf(v.w, 0.f) * m, ww.w, t4.w);
            }
            s_a[q * 128 + tid] = t4;
        }
        // residual maps of this pixel: in flight while the MMAs run
        const size_t base = (size_t)b * 8 * plane + (size_t)y * W + x;
        float4 rs[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            rs[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xin_ok) {
                const size_t off = base + q * plane;
                if (add_x) rs[q] = __ldg(reinterpret_cast<const float4*>(xin) + off);
                if (r1) rs[q] = f4_add(rs[q], ld_stream(reinterpret_cast<const float4*>(r1) + off));
                if (r2) rs[q] = f4_add(rs[q], ld_stream(reinterpret_cast<const float4*>(r2) + off));
            }
        }
        fence_proxy_async();                       // generic-proxy writes of s_a -> visible to the tensor core
        tc_fence_before();
        __syncthreads();                           // also: everyone finished reading TMEM of the previous row
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)
                tc_mma_tf32(tmem, a_desc0 + (uint64_t)(k8 * 2 * 128), b_desc0 + (uint64_t)(k8 * 64), tc_idesc(32u), k8 ? 1u : 0u);
            tc_commit(smem_u32(&bar));
        }
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1u;
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
        tc_fence_before();
        if (xin_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                reinterpret_cast<float4*>(out)[base + q * plane] =
                    make_float4(v[q * 4 + 0] + s_sh[q * 4 + 0] + rs[q].x, v[q * 4 + 1] + s_sh[q * 4 + 1] + rs[q].y,
                                v[q * 4 + 2] + s_sh[q * 4 + 2] + rs[q].z, v[q * 4 + 3] + s_sh[q * 4 + 3] + rs[q].w);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int dilconv_tc_launch(const float* x, const float* dw, const float* pw, const float* ch_scale, const float* ch_shift,
                      const float* r1, const float* r2, float* out, int add_x, int dil, int B, int H, int W,
                      cudaStream_t stream) {
    dim3 grid(cdiv(W, 128), cdiv(H, DC_ROWS), B);
    if (dil == 1) dilconv_tc_kernel<1><<<grid, 128, 0, stream>>>(x, dw, pw, ch_scale, ch_shift, r1, r2, out, add_x, H, W);
    else if (dil == 2) dilconv_tc_kernel<2><<<grid, 128, 0, stream>>>(x, dw, pw, ch_scale, ch_shift, r1, r2, out, add_x, H, W);
    else { set_error("dilconv_tc: dilation %d not instantiated", dil); return PAIF_ENOTSUP; }
    return check_launch("paif_dilconv_forward(tcgen05)");
}

bool conv_tc_supported(const PaifConvDesc& d) {
    if (d.cout != 32 || d.cin_per_src != 32 || d.kh != d.kw) return false;
    if (d.storage < 0 || d.storage > 2) return false;
    TcPlan p;
    return tc_make_plan(d.nsrc, d.kh, d.dil, d.storage == PAIF_STORAGE_BF16, &p);
}

static int tc_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

static int tc_rows_per_cta(const PaifConvDesc& d, const TcPlan& p) {
    // One CTA per SM at a time.  Pick the chunk height that minimises  waves x (rows + halo + prologue):
    // the prologue / drain of a CTA (barriers, TMEM, weight slab, pipeline fill) costs about as much as 8 rows,
    // and a last wave with a handful of CTAs costs a full wave (e.g. one 480x640 frame: 29 chunks of 17 rows =
    // 145 CTAs = one wave, where 8-row chunks would take three).  Multi-pass plans keep every output row of the
    // chunk in TMEM (<= 16 slots).  Measured at 16 x 480 x 640 (3x3 32->32): 30-row chunks 0.258 ms, 44..69 rows
    // 0.242 ms, 96 rows 0.247, 160 rows 0.263.
    const int strips = cdiv(d.W, TC_TW), sms = tc_num_sms();
    const int max_rch = p.npass > 1 ? TC_SLOTS : 128;    // single pass: TMEM slots are a ring, any chunk height works
    const int halo = 2 * p.pad;
    long long best_cost = -1;
    int best = max_rch;
    for (int rch = max_rch; rch >= 4; --rch) {
        const long long ctas = (long long)strips * cdiv(d.H, rch) * d.B;
        const long long waves = (ctas + sms - 1) / sms;
        const long long cost = waves * (rch + halo + 8);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rch; }
    }
    return best;
}

static int tc_persist_mode = -1;      // paif_conv_set_persistent / PAIF_TC_PERSIST

// Persistent launch shape: single-pass plans without per-CTA channel sums (those are summed per fixed tile so that
// the ECA statistics do not depend on the partition).  Returns the number of CTAs (0 = tiled launch).
static int tc_persistent_ctas(const TcPlan& p, bool partials, int B, int H, int W) {
    if (tc_persist_mode < 0) { const char* e = getenv("PAIF_TC_PERSIST"); tc_persist_mode = e ? atoi(e) : 1; }   // 0: tiled launches (A/B runs)
    if (!tc_persist_mode || p.npass != 1 || partials) return 0;
    const long long rows = (long long)B * cdiv(W, TC_TW) * H;
    if (rows > 0x7fffffffLL) return 0;
    const long long n = rows / 12;                         // at least ~12 rows per CTA: a prologue costs about 8
    return (int)(n < 1 ? 1 : (n > tc_num_sms() ? tc_num_sms() : n));
}

static void tc_set_grid(TcGeom& g, const TcPlan& p, bool partials, dim3* grid) {
    g.strips = cdiv(g.W, TC_TW);
    g.rows_total = g.B * g.strips * g.H;
    static int epf = -1;
    if (epf < 0) { const char* e = getenv("PAIF_TC_EPI_PREFETCH"); epf = e ? atoi(e) : 1; }       // 0: A/B runs
    g.epi_prefetch = epf;
    int n = tc_persistent_ctas(p, partials, g.B, g.H, g.W);
    g.persist = n > 0;
    static int banded = -1;
    if (banded < 0) { const char* e = getenv("PAIF_TC_BANDS"); banded = e ? atoi(e) : 1; }       // 0: linear shares (A/B runs)
    // (the tensor-bound shapes — 7x7, the 5x5 single-output stencil — keep the linear form: every SM busy is worth more
    //  to them than L2 hits; 148 vs 145 CTAs at 5 strips)
    if (n > 0 && banded && g.strips <= n && g.k < 5) {
        long long nb = n / g.strips;                            // bands of image rows; `strips` CTAs per band
        const long long most = ((long long)g.B * g.H) / 12;
        if (nb > most) nb = most < 1 ? 1 : most;
        n = (int)nb * g.strips;
        g.persist = 2;
    }
    *grid = n > 0 ? dim3(n, 1, 1) : dim3(g.strips, cdiv(g.H, g.RCH), g.B);
}

int conv_tc_tiles(int H, int W) {
    // upper bound independent of the per-launch row chunk: the smallest chunk is 4 rows
    return cdiv(W, TC_TW) * cdiv(H, 4);
}

int conv_tc_launch(const PaifConvDesc& d, cudaStream_t stream) {
    TcGeom g;
    if (!tc_make_plan(d.nsrc, d.kh, d.dil, d.storage == PAIF_STORAGE_BF16, &g.plan)) { set_error("conv_tc: no plan"); return PAIF_ENOTSUP; }
    if (d.storage != PAIF_STORAGE_F32 && (d.pre_res[0] || d.pre_res[1]) && d.post_res[2] && d.pre_res[1]) {
        // the bf16 epilogue keeps the fetched residual rows packed in four register slots
        set_error("conv_tc: bf16 storage takes at most 4 residual maps (3 post + 2 pre given)");
        return PAIF_ENOTSUP;
    }
    g.B = d.B; g.H = d.H; g.W = d.W; g.nsrc = d.nsrc; g.k = d.kh; g.dil = d.dil;
    g.RCH = tc_rows_per_cta(d, g.plan);
    g.tiles_alloc = conv_tc_tiles(d.H, d.W);
    for (int i = 0; i < 3; ++i) g.src[i] = d.src[i];
    g.wmma = d.weight_mma;
    EpiParams e = make_epi(d);
    dim3 grid;
    tc_set_grid(g, g.plan, d.chan_partials != nullptr, &grid);
    if (d.chan_partials) {
        // partial-sum slots beyond this launch's tile count must read as zero
        cudaError_t err = cudaMemsetAsync(d.chan_partials, 0, (size_t)d.B * conv_tc_tiles(d.H, d.W) * 32 * sizeof(float), stream);
        if (err != cudaSuccess) { set_error("conv_tc memset: %s", cudaGetErrorString(err)); return (int)err; }
    }
    const int kk = d.kh == 1 ? 1 : d.kh, dd = d.kh == 1 ? 1 : d.dil;     // a 1x1 kernel has no dilation
#define TC_CASE(K_, D_, Q_, S_)                                                                                    \
    if (kk == K_ && dd == D_ && g.plan.KQ == Q_ && d.storage == S_) {                                              \
        static unsigned long long attr_done = 0;                                                                    \
        int dev_;                                                                                                   \
        if (attr_needed(attr_done, &dev_)) {                                                                        \
            cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel<K_, D_, Q_, false, S_>,                           \
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET); \
            if (err == cudaSuccess)                                                                                 \
                err = cudaFuncSetAttribute(conv_tc_kernel<K_, D_, Q_, true, S_>,                                    \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET);     \
            if (err != cudaSuccess) { set_error("conv_tc smem attr: %s", cudaGetErrorString(err)); return (int)err; } \
            attr_mark(attr_done, dev_);                                                                             \
        }                                                                                                           \
        if (d.chan_partials) conv_tc_kernel<K_, D_, Q_, true, S_><<<grid, TC_NT, g.plan.smem_bytes, stream>>>(g, e); \
        else conv_tc_kernel<K_, D_, Q_, false, S_><<<grid, TC_NT, g.plan.smem_bytes, stream>>>(g, e);              \
        return check_launch("paif_conv_forward(tcgen05)");                                                          \
    }
    TC_CASE(1, 1, 8, 0) TC_CASE(3, 1, 8, 0) TC_CASE(3, 2, 8, 0) TC_CASE(5, 1, 8, 0) TC_CASE(5, 2, 8, 0)
    TC_CASE(7, 1, 4, 0) TC_CASE(7, 2, 4, 0)
    // bf16 storage: a whole 32-channel source is 4 planes, so every shape runs with KQ = 4 (7x7 in a single pass)
    TC_CASE(1, 1, 4, 1) TC_CASE(3, 1, 4, 1) TC_CASE(3, 2, 4, 1) TC_CASE(5, 1, 4, 1) TC_CASE(5, 2, 4, 1)
    TC_CASE(7, 1, 4, 1) TC_CASE(7, 2, 4, 1)
    // fp32 sources, bf16 outputs: the 1x1 convolutions that enter the bf16 part of the network
    TC_CASE(1, 1, 8, 2)
#undef TC_CASE
    set_error("conv_tc: no kernel instance for k=%d dil=%d KQ=%d storage=%d", d.kh, d.dil, g.plan.KQ, d.storage);
    return PAIF_ENOTSUP;
}

// stem_out's merged 5x5 32->1 stencil (interior-class weights) + PReLU + tanh on the engine: single-output mode (CP = 16).
// feat: fp32 C4 map, or bf16 C8 map when bf16 != 0; wmma: the 5x5 weights packed like any other weight image with the
// output channel padded to 16 (channel 0 real).  The one-pixel image border is redone by out_border_kernel afterwards.
int out_tc_launch(const void* feat, const void* wmma, const float* slope, float* out, float* pre_out,
                  int bf16, int B, int H, int W, cudaStream_t stream) {
    TcGeom g;
    if (!tc_make_plan(1, 5, 1, bf16 != 0, &g.plan, 16)) { set_error("out_tc: no plan"); return PAIF_ENOTSUP; }
    PaifConvDesc d = {};
    d.B = B; d.H = H; d.W = W;
    g.B = B; g.H = H; g.W = W; g.nsrc = 1; g.k = 5; g.dil = 1;
    g.RCH = tc_rows_per_cta(d, g.plan);
    g.tiles_alloc = 0;
    g.src[0] = feat; g.src[1] = g.src[2] = nullptr;
    g.wmma = wmma;
    EpiParams e = {};
    e.slope = slope; e.post_scale = 1.f; e.out = out; e.out_pre = pre_out; e.H = H; e.W = W;
    dim3 grid;
    tc_set_grid(g, g.plan, false, &grid);
    static unsigned long long attr_done = 0;
    int dev;
    if (attr_needed(attr_done, &dev)) {
        cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel<5, 1, 8, false, 0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(conv_tc_kernel<5, 1, 4, false, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET);
        if (err != cudaSuccess) { set_error("out_tc smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        attr_mark(attr_done, dev);
    }
    if (bf16) conv_tc_kernel<5, 1, 4, false, 1, 16><<<grid, TC_NT, g.plan.smem_bytes, stream>>>(g, e);
    else conv_tc_kernel<5, 1, 8, false, 0, 16><<<grid, TC_NT, g.plan.smem_bytes, stream>>>(g, e);
    return check_launch("paif_out_forward_tc");
}

#ifdef PAIF_TC_PROFILE
extern "C" int paif_debug_tc_counters(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, tc_prof, sizeof(tc_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tc_prof, z, sizeof(z)); }
    return 0;
}
#endif

}  // namespace paif

extern "C" int paif_conv_set_persistent(int on) {
    if (paif::tc_persist_mode < 0) { const char* e = getenv("PAIF_TC_PERSIST"); paif::tc_persist_mode = e ? atoi(e) : 1; }
    const int prev = paif::tc_persist_mode;
    paif::tc_persist_mode = on ? 1 : 0;
    return prev;
}

namespace paif {

int conv_tc_kq(int nsrc, int k, int dil, bool bf16) {
    TcPlan p;
    return tc_make_plan(nsrc, k, dil, bf16, &p) ? p.KQ : 0;
}

}  // namespace paif
