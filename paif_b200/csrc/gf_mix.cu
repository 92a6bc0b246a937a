// Fused guided-filter decomposition + folded 128->32 1x1 (Cell_Decom.decomposition + conv1x1_lf/hf,
// core/model_fusion_auto.py:509-535): one kernel emits the branch input
//     x = Wa LF_1e-3 + Wb LF_1e-4 + Wc z + bias          (Wa/Wb/Wc: the folded 1x1, fusion.py::_fold_decomp_1x1)
// without materialising the LF maps.  It uses the linearity of everything after the level-1 statistics:
//     LF_e,c = mean2(A_e,c) g + mean2(b_e,c),  A_e,c = cov_c inv_e,  b_e,c = mz_c - A_e,c mx
//  => sum_c Wa[o,c] LF_1,c + Wb[o,c] LF_2,c = mean2(A'_o) g + mean2(b'_o)  with
//     A'_o = inv_1 (Wa cov)_o + inv_2 (Wb cov)_o,   b'_o = ((Wa+Wb) mz)_o - mx A'_o
// so the channel mix happens BETWEEN the two box-filter levels (64 level-2 box filters instead of 128) and runs on
// the tensor cores (tcgen05, TF32 operands, fp32 accumulate): per level-1 row three small GEMMs
//     [P|Q] = cov [Wa|Wb]^T (K=32, N=64),  R = mz (Wa+Wb)^T (K=32, N=32),  C = z[yo] Wc^T (K=32, N=32).
//
// One persistent CTA per SM walks (image, 48-column strip, row chunk) work items.  Roles (14 warps):
//   warps 0-3  L1: level-1 box filters in the marching layout of gf.cu (half-warp = one channel quad x 64 raw columns,
//              lane = 4 columns; vertical running sums in registers, horizontal 9-sums by shuffles); writes cov / mz of
//              two level-1 rows as one M=128 UMMA A operand (K-major, no swizzle) into shared memory;
//   warp 12    MMA issuer (one thread) + TMEM allocator;  warp 13  producer: cp.async.bulk of the raw z rows of the
//              output rows (A operand of the Wc GEMM);
//   warps 8-11 EP: TMEM -> registers (thread = pixel), A' = inv1 P + inv2 Q, b' = R - mx A', C + bias, written to a
//              shared-memory exchange buffer in quad-plane order (XOR-swizzled so that both sides are conflict-free);
//   warps 4-7  L2: level-2 box filters in the marching layout (half-warp = one OUTPUT channel quad); the 9-row history
//              of (A', b') that the vertical running sum needs lives in TMEM (tcgen05.st / ld, 32 columns per row and
//              warp: 9 x 32 columns next to the 128 accumulator columns); output = mean2(A') g + mean2(b') + C.
// The level-2 history is what bounds the strip width: 9 rows x 64 values x 4 B = 2.3 KB per pixel column.
// Details that carry the performance ("second pass"): the level-1 ring reads are rotated per lane so that
// they are bank-conflict free; columns outside the image read zeros (a zero page in global memory for the leaving row,
// ring columns cleared by the producer for the entering row) instead of being masked; every address is a 32-bit offset
// from a per-chunk base; thread index and shared-window base are kept out of the row loops; a ring stage / exchange row
// is released by an arrive whose ADDRESS depends on the loaded registers (mbar_arrive_after_loads).
// SAVE_MA (paif_gf_mix_forward_save): also writes mean2(A'), the direct guide term of the adjoint (gf.cu).
#ifdef PAIF_SANITIZER_BUILD
#define PAIF_MBAR_SPIN_LIMIT 0xffffffffu        // compute-sanitizer slows the busy roles by orders of magnitude: no poll-count watchdog
#else
#define PAIF_MBAR_SPIN_LIMIT (1u << 25)        // a lost arrival traps after ~10 s of polling
#endif
#define PAIF_MBAR_QUIET
#ifndef PAIF_NO_SUSPEND_HINT                   // (builds for compute-sanitizer runs define it: plain polling)
#define PAIF_MBAR_SUSPEND_NS 20000
#endif
#include <cstddef>
#include "tc_ptx.cuh"

namespace paif {
namespace gx {

#ifdef PAIF_TC_PROFILE
// development-only role timeline: wait / busy cycles summed over all CTAs (read with paif_debug_gx_counters)
__device__ unsigned long long gx_prof[16];
#define GX_PROF_DECL long long prof_wait = 0, prof_t0 = clock64()
#define GX_WAIT(stmt) do { const long long t_ = clock64(); stmt; prof_wait += clock64() - t_; } while (0)
#define GX_PROF_END(slot) do { if ((threadIdx.x & 31) == 0) { atomicAdd(&gx_prof[(slot) * 2], (unsigned long long)prof_wait); \
                                   atomicAdd(&gx_prof[(slot) * 2 + 1], (unsigned long long)(clock64() - prof_t0)); } } while (0)
#else
#define GX_PROF_DECL
#define GX_WAIT(stmt) stmt
#define GX_PROF_END(slot)
#endif

__device__ __align__(64) const float gx_zero[16] = {0.f};      // what out-of-image lanes and absent leaving rows load
constexpr int OUTW = 48;                       // output columns per strip (64 raw -> 56 level-1 -> 48 output)
constexpr int NWARPS = 14, NT = NWARPS * 32;
constexpr int MMA_WARP = 12;            // warp 13: producer
constexpr int AOP_SBO = 144;                   // core-matrix pitch of the A operand (128 B + 16 B pad: conflict-free 4-column stores)
constexpr int AOP_PLANE = 16 * AOP_SBO;        // one 16-byte K chunk (4 channels) x 128 rows
constexpr int AOP_BYTES = 16 * AOP_PLANE;      // planes 0-7: cov quads, 8-15: mean_z quads
constexpr int X_PLANE = 1024;                  // exchange, one buffer per row of the pair: [plane][64 slots][16 B]
constexpr int X_BYTES = 24 * X_PLANE;          // planes 0-7 A', 8-15 b', 16-23 C
constexpr int Z_BYTES = 8 * 2048;              // 8 quad planes x 128 rows x 16 B (rows 0-63: first output row, 64-127: second)
constexpr int W_BYTES = 16384;                 // [Wa|Wb] 8 KB, Wa+Wb 4 KB, Wc 4 KB (TF32, UMMA B tiles)
constexpr int RS = 6;                          // raw-row ring stages (entering rows of the level-1 window, fetched by the producer)
constexpr int R_BYTES = 8 * 1024 + 256;        // 8 quad planes x 64 pixels x 16 B + 64 guide values
constexpr int OFF_W = 0, OFF_AOP = OFF_W + W_BYTES, OFF_X = OFF_AOP + 2 * AOP_BYTES, OFF_Z = OFF_X + 2 * X_BYTES,
              OFF_R = OFF_Z + 2 * Z_BYTES, OFF_BARS = OFF_R + RS * R_BYTES, SMEM_BYTES = OFF_BARS + 1024;
constexpr int RING_COL0 = 128, RING_SLOTS = 10;     // 9 rows of history + the row being written (so that the old row is read from another slot)

struct Bars {
    uint64_t aop_full[2], aop_empty[2], z_full[2], x_full[2], x_empty[2], d_full, d_empty, r_full[RS], r_empty[RS];
    uint32_t tmem_base;
    alignas(16) float bias[32];
};
static_assert(sizeof(Bars) <= 1024, "barrier block");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
    const float* feat; const float* guide; const float* stats; const void* wpack; const float* bias;
    void* out;
    float* mean_a;             // SAVE_MA: mean2(A') of every output pixel, fp32 C4 map (the adjoint's direct guide term)
    int B, H, W, nstrips;
    int RC, nchunks;           // rows per chunk / chunks per strip: a function of the shape only (see the launcher)
    int nitems;                // B * nstrips * nchunks work items, split into contiguous ranges over the CTAs
};

__device__ __forceinline__ float win_count(int p, int n) {
    const int lo = p - 4 < 0 ? 0 : p - 4, hi = p + 4 > n - 1 ? n - 1 : p + 4;
    return (float)(hi - lo + 1);
}
// 1 / win_count(p, n) for n > 9 (5..9 rows or columns in the window): the correctly rounded constants __frcp_rn returns
__device__ __forceinline__ float rcp_count(int p, int n) {
    const int lo = p - 4 < 0 ? 0 : p - 4, hi = p + 4 > n - 1 ? n - 1 : p + 4, c = hi - lo + 1;
    return c == 9 ? 1.f / 9.f : c == 8 ? 1.f / 8.f : c == 7 ? 1.f / 7.f : c == 6 ? 1.f / 6.f : 1.f / 5.f;
}
// o[k] = sum of columns (4j+k) .. (4j+k+8) of the per-lane column quadruples a[0..3]; 16-lane segments
__device__ __forceinline__ void hsum9(const float (&a)[4], float (&o)[4]) {
    const float p01 = a[0] + a[1], p23 = a[2] + a[3], p012 = p01 + a[2], full = p01 + p23;
    const float n1 = __shfl_down_sync(0xffffffffu, full, 1, 16);
    const float v8 = __shfl_down_sync(0xffffffffu, a[0], 2, 16);
    const float v89 = __shfl_down_sync(0xffffffffu, p01, 2, 16);
    const float v8a = __shfl_down_sync(0xffffffffu, p012, 2, 16);
    const float v8b = __shfl_down_sync(0xffffffffu, full, 2, 16);
    o[0] = full + n1 + v8;
    o[1] = (a[1] + p23) + n1 + v89;
    o[2] = p23 + n1 + v8a;
    o[3] = a[3] + n1 + v8b;
}
// the same on channel pairs: FADD2 halves the adds (the kernel is issue-bound, not FP32-pipe-bound); shuffles move 32 bits
__device__ __forceinline__ float2 shfl_down2(float2 v, int d) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, d, 16), __shfl_down_sync(0xffffffffu, v.y, d, 16));
}
__device__ __forceinline__ void hsum9x2(const float2 (&a)[4], float2 (&o)[4]) {
    const float2 p01 = __fadd2_rn(a[0], a[1]), p23 = __fadd2_rn(a[2], a[3]);
    const float2 p012 = __fadd2_rn(p01, a[2]), full = __fadd2_rn(p01, p23);
    const float2 n1 = shfl_down2(full, 1), v8 = shfl_down2(a[0], 2), v89 = shfl_down2(p01, 2);
    const float2 v8a = shfl_down2(p012, 2), v8b = shfl_down2(full, 2);
    o[0] = __fadd2_rn(__fadd2_rn(full, n1), v8);
    o[1] = __fadd2_rn(__fadd2_rn(__fadd2_rn(a[1], p23), n1), v89);
    o[2] = __fadd2_rn(__fadd2_rn(p23, n1), v8a);
    o[3] = __fadd2_rn(__fadd2_rn(a[3], n1), v8b);
}
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }
__device__ __forceinline__ void hsum9c(const float (&s)[4][4], int c, float (&o)[4]) {
    const float a[4] = {s[0][c], s[1][c], s[2][c], s[3][c]};
    hsum9(a, o);
}
// 4 consecutive floats of a plane row (x multiple of 4, W multiple of 4): zero outside [0, W)
__device__ __forceinline__ void ld_cols4(const float* __restrict__ row, int x, int W, float (&v)[4]) {
    // branch-free: the address is clamped into the row, the value selected afterwards (no divergence around the shuffles)
    const bool ok = x >= 0 && x < W;
    const float4 t = __ldg(reinterpret_cast<const float4*>(row + (ok ? x : 0)));
    v[0] = ok ? t.x : 0.f; v[1] = ok ? t.y : 0.f; v[2] = ok ? t.z : 0.f; v[3] = ok ? t.w : 0.f;
}
// the 4 channels of one quad at pixels x..x+3 of a C4 map row (`row` -> (quad plane, row y, pixel 0)); zero outside
__device__ __forceinline__ void ld_quad(const float* __restrict__ row, int x, int W, float (&z)[4][4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        const bool ok = xx >= 0 && xx < W;
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + (size_t)(ok ? xx : 0) * 4));
        z[k][0] = ok ? t.x : 0.f; z[k][1] = ok ? t.y : 0.f; z[k][2] = ok ? t.z : 0.f; z[k][3] = ok ? t.w : 0.f;
    }
}
// Unmasked variants: the address is clamped into the row and the caller applies the column mask where the value is USED
// (one iteration later) — a select placed next to the load would make the warp wait for the load right there.
__device__ __forceinline__ float4 ld_cols4_raw(const float* __restrict__ row, int x, int W) {
    return __ldg(reinterpret_cast<const float4*>(row + ((x >= 0 && x < W) ? x : 0)));
}
__device__ __forceinline__ void ld_quad_raw(const float* __restrict__ row, int x, int W, float4 (&z)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        z[k] = __ldg(reinterpret_cast<const float4*>(row + (size_t)((xx >= 0 && xx < W) ? xx : 0) * 4));
    }
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
        "tcgen05.wait::st.sync.aligned;"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}

__device__ __forceinline__ void tmem_st32_nowait(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
// tcgen05.ld without the wait: the caller issues tcgen05.wait::ld before it reads v[]
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// completes a tmem_ld32_nowait: the registers are operands of the wait so that no use of v[] can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld32(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                   "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]),
                   "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                   "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :: "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n\t"
        "tcgen05.wait::st.sync.aligned;"          // in the same statement: the source registers are read asynchronously
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld16(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                   "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :: "memory");
}

// shared-memory accesses by 32-bit shared-window address (the level-1 role: no generic-to-shared conversion per access)
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// mbarrier arrive that releases a shared-memory stage the calling thread has just READ with ordinary loads.  The arrive
// must not be issued while a load is still in flight (the producer would refill the stage under it), and an inline-asm
// operand that the instruction does not use is no dependency for ptxas.  So the barrier ADDRESS is made to depend on
// one register of every load: (bits of the loaded values) & rt_zero, where rt_zero is a run-time zero the compiler
// cannot see through (blockIdx.y of a 1-D grid).
__device__ __forceinline__ void mbar_arrive_after_loads(uint32_t bar, uint32_t rt_zero, float a, float b, float c, float d, float e) {
    const uint32_t dep = (__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c) | __float_as_uint(d) | __float_as_uint(e)) & rt_zero;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + dep) : "memory");
}

// one work chunk: rows [y0, y0 + rows) of strip `strip` of image b
struct Chunk { int b, strip, y0, rows; };
struct Walker {
    int t, t1, H, nstrips, RC, nchunks;
    __device__ __forceinline__ Walker(const Params& p) : H(p.H), nstrips(p.nstrips), RC(p.RC), nchunks(p.nchunks) {
        t = (int)((long long)p.nitems * blockIdx.x / gridDim.x);
        t1 = (int)((long long)p.nitems * (blockIdx.x + 1) / gridDim.x);
    }
    __device__ __forceinline__ bool next(Chunk& c) {
        if (t >= t1) return false;
        int item = t++;
        const int chunk = item % nchunks; item /= nchunks;
        c.strip = item % nstrips;
        c.b = item / nstrips;
        c.y0 = chunk * RC;
        c.rows = min(RC, H - c.y0);
        return true;
    }
};

// exchange-buffer slot (16-byte units inside one plane-row of 64) of level-1 / output column index p = 4j + k:
// the 4 columns of a lane are 16 slots apart, XOR-swizzled so that both the thread-per-pixel writers (8 consecutive
// p per phase) and the 4-columns-per-lane readers (8 consecutive j per phase) touch 8 distinct bank groups
__device__ __forceinline__ int xslot(int j, int k) { return (k << 4) + (j ^ (k << 1)); }

template <bool OUT_BF, bool SAVE_MA>
__global__ void __launch_bounds__(NT, 1)    // 14 warps = up to 4 per SM sub-partition (16 K registers each): 128 registers per thread
gf_mix_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BARS);
    // threadIdx.x and the shared-memory window base through volatile asm: the compiler cannot re-read the special
    // registers inside the row loops (S2R / S2UR have ~25 cycles of latency, and the level-1 warp is latency-bound)
    uint32_t tid_op, sbase;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_op));
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
    const int tid = (int)tid_op, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler: the role branches do not diverge
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->aop_full[i]), 128);
            mbar_init(smem_u32(&bars->aop_empty[i]), 1);
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->x_full[i]), 64);        // the 64 EP threads of that row of the pair
            mbar_init(smem_u32(&bars->x_empty[i]), 128);
        }
        for (int i = 0; i < RS; ++i) { mbar_init(smem_u32(&bars->r_full[i]), 1); mbar_init(smem_u32(&bars->r_empty[i]), 128); }
        mbar_init(smem_u32(&bars->d_full), 1);
        mbar_init(smem_u32(&bars->d_empty), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) bars->bias[tid] = p.bias ? p.bias[tid] : 0.f;
    for (int i = tid; i < W_BYTES / 16; i += NT)
        reinterpret_cast<uint4*>(smem + OFF_W)[i] = __ldg(reinterpret_cast<const uint4*>(p.wpack) + i);
    for (int i = tid; i < RS * R_BYTES / 16; i += NT)      // ring columns outside the image are never written by the
        reinterpret_cast<uint4*>(smem + OFF_R)[i] = make_uint4(0u, 0u, 0u, 0u);    // copies and are masked by 0 * value: keep them finite
    fence_proxy_async();                                   // the weight tiles are read by the tensor core (async proxy)
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

    Walker walk(p);
    Chunk ck;
    uint32_t gp0 = 0;                                      // pairs of level-1 rows processed so far (all roles agree)

    if (warp < 4) {
        // ============================ L1: level-1 statistics -> A operand ============================
        const int h = lane >> 4, j = lane & 15, q = 2 * warp + h;
        const int rot = (j >> 1) & 3;
        const uint32_t rt_zero = blockIdx.y;                   // 0 at run time, opaque to the compiler (mbar_arrive_after_loads)
        const uint32_t aop = sbase + OFF_AOP;
        // byte offset of row m = half*64 + 4j + k inside a plane: (m >> 3) * SBO + (m & 7) * 16
        const int aoff = (j >> 1) * AOP_SBO + (j & 1) * 64;
        const uint32_t ring = sbase + OFF_R;
        uint32_t rcount = 0;                                   // raw rows taken from the ring so far
        GX_PROF_DECL;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xr = x0 - 8 + 4 * j, xs = xr + 4;
            // x0, xr, xs and W are multiples of 4: the 4 raw columns of a lane are all inside the image or all outside.
            // Out-of-image lanes (and leaving rows that do not exist) read a 64-byte page of zeros instead of being
            // masked value by value, and every address is one 32-bit element offset from a per-lane base.
            const bool cin = xr >= 0 && xr < W;
            const float4* zcol = reinterpret_cast<const float4*>(p.feat) + ((size_t)ck.b * 8 + q) * plane + (cin ? xr : 0);
            const float* gcol = p.guide + (size_t)ck.b * plane + (cin ? xr : 0);
            const float* mcol = p.stats + (size_t)ck.b * plane + ((xs >= 0 && xs < W) ? xs : 0);
            const float4* zero4 = reinterpret_cast<const float4*>(gx_zero);
            float cs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                cs[k] = (xs + k >= 0 && xs + k < W && 4 * j + k < 56) ? __frcp_rn(win_count(xs + k, W)) : 0.f;
            float Sz[4][4], Sgz[4][4];
            float4 zq4[4], gq4, mx4;                               // leaving row (zeros where it does not exist) / mean_g of the next iteration
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                zq4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) Sz[k][c] = Sgz[k][c] = 0.f;
            }
            gq4 = mx4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const int n1 = rows + 8, nt = rows + 16;
            for (int t = 0; t < nt; ++t) {
                const int yr = y0 - 8 + t;
                // ---- leaving row (fetched from global memory one iteration ago; zeros where there is none)
                const float zq[4][4] = {{zq4[0].x, zq4[0].y, zq4[0].z, zq4[0].w}, {zq4[1].x, zq4[1].y, zq4[1].z, zq4[1].w},
                                        {zq4[2].x, zq4[2].y, zq4[2].z, zq4[2].w}, {zq4[3].x, zq4[3].y, zq4[3].z, zq4[3].w}};
                const float gq[4] = {gq4.x, gq4.y, gq4.z, gq4.w};
                const float mxc[4] = {mx4.x, mx4.y, mx4.z, mx4.w};     // mean_g of the level-1 row this iteration completes
                // ---- entering row: from the shared-memory ring the producer fills a few rows ahead
                float zn[4][4], gn[4];
                {
                    const uint32_t st = rcount % RS;
                    GX_WAIT(mbar_wait((sbase + OFF_BARS + (uint32_t)offsetof(Bars, r_full) + 8u * (st)), (rcount / RS) & 1u));
                    if (yr >= 0 && yr < H) {
                        const uint32_t rs_ = ring + st * R_BYTES;
                        const float4 g4 = lds128(rs_ + 8192 + j * 16);
                        // ring columns outside the image hold zeros (the producer clears them at the start of the chunk).
                        // A lane's 4 pixels are 64 contiguous bytes, so "pixel k of every lane" is a 4-way bank conflict:
                        // the k-th load of lane j takes pixel (k + j/2) % 4 instead (8 consecutive lanes then cover the 8
                        // 16-byte slots of a 128-byte line) and two select stages put the pixels back in order.
                        gn[0] = g4.x; gn[1] = g4.y; gn[2] = g4.z; gn[3] = g4.w;
                        float4 v[4], w[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            v[k] = lds128(rs_ + q * 1024 + j * 64 + (((k + rot) & 3) << 4));
                        const bool r1b = rot & 1, r2b = rot & 2;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 x = v[k], y = v[(k + 3) & 3];
                            w[k] = make_float4(r1b ? y.x : x.x, r1b ? y.y : x.y, r1b ? y.z : x.z, r1b ? y.w : x.w);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 x = w[k], y = w[(k + 2) & 3];
                            zn[k][0] = r2b ? y.x : x.x; zn[k][1] = r2b ? y.y : x.y; zn[k][2] = r2b ? y.z : x.z; zn[k][3] = r2b ? y.w : x.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { gn[k] = 0.f; zn[k][0] = zn[k][1] = zn[k][2] = zn[k][3] = 0.f; }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            Sz[k][c] += zn[k][c] - zq[k][c];
                            Sgz[k][c] = __fmaf_rn(-gq[k], zq[k][c], __fmaf_rn(gn[k], zn[k][c], Sgz[k][c]));
                        }
                    mbar_arrive_after_loads((sbase + OFF_BARS + (uint32_t)offsetof(Bars, r_empty) + 8u * (st)), rt_zero, gn[0], zn[0][3], zn[1][3], zn[2][3], zn[3][3]);
                    ++rcount;
                }
                {   // global loads for iteration t+1, issued once this iteration's leaving row has been consumed (the registers are
                    // reused without copies; most of an iteration is still ahead to cover the L2 latency): leaving row yr-8 (an L2
                    // hit: it entered 9 rows ago), mean_g of row yr-3
                    // (row clamped into the image: whether it counts is decided by the window count at the use)
                    const int yl = yr - 8, ysn = yr - 3;
                    const bool lv = t + 1 < nt && t + 1 >= 9 && yl >= 0 && yl < H && cin;
                    const unsigned rowl = (unsigned)(min(max(yl, 0), H - 1) * W), rowm = (unsigned)(min(max(ysn, 0), H - 1) * W);
                    const float4* zl = lv ? zcol + rowl : zero4;
                    const float4* gl = lv ? reinterpret_cast<const float4*>(gcol + rowl) : zero4;
                    zq4[0] = __ldg(zl); zq4[1] = __ldg(zl + 1); zq4[2] = __ldg(zl + 2); zq4[3] = __ldg(zl + 3);
                    gq4 = __ldg(gl);
                    mx4 = __ldg(reinterpret_cast<const float4*>(mcol + rowm));
                }
                if (t < 8) continue;
                const int r1 = t - 8, ys = yr - 4, half = r1 & 1;
                const uint32_t gpair = gp0 + (uint32_t)(r1 >> 1), buf = gpair & 1u;
                float rn[4];
                {
                    const float rcy = (ys >= 0 && ys < H) ? rcp_count(ys, H) : 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) rn[k] = rcy * cs[k];
                }
                float mz[4][4], cov[4][4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float bz[4], bg[4];
                    hsum9c(Sz, c, bz);
                    hsum9c(Sgz, c, bg);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mz[k][c] = bz[k] * rn[k];
                        cov[k][c] = __fmaf_rn(-mxc[k], mz[k][c], bg[k] * rn[k]);
                    }
                }
                if (half == 0) GX_WAIT(mbar_wait((sbase + OFF_BARS + (uint32_t)offsetof(Bars, aop_empty) + 8u * (buf)), ((gpair >> 1) & 1u) ^ 1u));
                const uint32_t base = aop + buf * AOP_BYTES + half * (8 * AOP_SBO) + aoff;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    sts128(base + q * AOP_PLANE + k * 16, cov[k][0], cov[k][1], cov[k][2], cov[k][3]);
                    sts128(base + (8 + q) * AOP_PLANE + k * 16, mz[k][0], mz[k][1], mz[k][2], mz[k][3]);
                }
                if (half == 1 || r1 == n1 - 1) {
                    fence_proxy_async();
                    mbar_arrive((sbase + OFF_BARS + (uint32_t)offsetof(Bars, aop_full) + 8u * (buf)));
                }
            }
            gp0 += (uint32_t)((n1 + 1) >> 1);
        }
        GX_PROF_END(0);
    } else if (warp < 8) {
        // ============================ L2: level-2 box filters -> output ============================
        const int w2 = warp & 3, h = lane >> 4, j = lane & 15, q = 2 * w2 + h;
        const uint32_t rt_zero = blockIdx.y;
        const uint32_t xbuf = sbase + OFF_X;
        const uint32_t ring = tmem_base + ((uint32_t)(w2 * 32) << 16) + RING_COL0;
        int xo_slot[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xo_slot[k] = xslot(j, k) * 16;
        GX_PROF_DECL;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xo = x0 + 4 * j;
            // per-chunk bases; inside the row loop every address is base + one 32-bit row offset
            const float* gcol = p.guide + (size_t)ck.b * plane + ((xo >= 0 && xo < W) ? xo : 0);
            float4* const o32 = reinterpret_cast<float4*>(p.out) + ((size_t)ck.b * 8 + q) * plane + xo;
            uint4* const o16 = reinterpret_cast<uint4*>(p.out) + ((size_t)ck.b * 4 + w2) * plane + xo;
            float4* const oma = reinterpret_cast<float4*>(p.mean_a) + ((size_t)ck.b * 8 + q) * plane + xo;
            float co[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                co[k] = (4 * j + k < OUTW && xo + k < W) ? __frcp_rn(win_count(xo + k, W)) : 0.f;
            float2 SA[4][2], SB[4][2];                             // vertical running sums of A' and b': [column k][channel pair]
#pragma unroll
            for (int k = 0; k < 4; ++k) SA[k][0] = SA[k][1] = SB[k][0] = SB[k][1] = make_float2(0.f, 0.f);
            const int n1 = rows + 8;
            int slot = 0;
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);           // guide at the next output row (fetched one row ahead)
            for (int r1 = 0; r1 < ((n1 + 1) & ~1); ++r1) {
                const int half = r1 & 1;
                const uint32_t gpair = gp0 + (uint32_t)(r1 >> 1);
                if (r1 >= n1) {
                    // odd row count: the second row of the last pair does not exist, but EP hands over both rows of
                    // every pair — take it and give it back so that the barrier phases stay in step
                    GX_WAIT(mbar_wait((sbase + OFF_BARS + (uint32_t)offsetof(Bars, x_full) + 8u * (half)), gpair & 1u));
                    mbar_arrive((sbase + OFF_BARS + (uint32_t)offsetof(Bars, x_empty) + 8u * (half)));
                    continue;
                }
                GX_WAIT(mbar_wait((sbase + OFF_BARS + (uint32_t)offsetof(Bars, x_full) + 8u * (half)), gpair & 1u));
                const uint32_t xrow = xbuf + half * X_BYTES;
                const bool has_old = r1 >= 9;
                const uint32_t t_new = ring + slot * 32, t_old = ring + (slot == RING_SLOTS - 1 ? 0 : slot + 1) * 32;   // row r1 - 9
                // the two halves (A', b') one after the other: new row from the exchange buffer, into the history ring and
                // into the running sums; the row leaving the window out of the ring and out of the sums
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float nw[16], od[16];
                    if (has_old) tmem_ld16_nowait(t_old + e * 16, od);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 a = lds128(xrow + (e * 8 + q) * X_PLANE + xo_slot[k]);
                        nw[k * 4 + 0] = a.x; nw[k * 4 + 1] = a.y; nw[k * 4 + 2] = a.z; nw[k * 4 + 3] = a.w;
                    }
                    tmem_st16(t_new + e * 16, nw);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const float2 v = make_float2(nw[k * 4 + 2 * h2], nw[k * 4 + 2 * h2 + 1]);
                            if (e == 0) SA[k][h2] = __fadd2_rn(SA[k][h2], v); else SB[k][h2] = __fadd2_rn(SB[k][h2], v);
                        }
                    if (has_old) {
                        tmem_wait_ld16(od);
                        const float2 neg = make_float2(-1.f, -1.f);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int h2 = 0; h2 < 2; ++h2) {
                                const float2 v = make_float2(od[k * 4 + 2 * h2], od[k * 4 + 2 * h2 + 1]);
                                if (e == 0) SA[k][h2] = __ffma2_rn(v, neg, SA[k][h2]); else SB[k][h2] = __ffma2_rn(v, neg, SB[k][h2]);
                            }
                    }
                }
                slot = slot == RING_SLOTS - 1 ? 0 : slot + 1;
                const float gc[4] = {g4.x, g4.y, g4.z, g4.w};
                float cdep[4] = {0.f, 0.f, 0.f, 0.f};
                {   // guide of the NEXT output row (row clamped into the chunk, columns into the image: only stored pixels
                    // use it): in flight during this row's horizontal sums
                    const int yon = min(max(y0 + r1 - 7, y0), y0 + rows - 1);
                    g4 = __ldg(reinterpret_cast<const float4*>(gcol + (unsigned)(yon * W)));
                }
                if (r1 >= 8) {
                    const int yo = y0 + r1 - 8;
                    float rno[4];
                    const float rcy = __frcp_rn(win_count(yo, H));
#pragma unroll
                    for (int k = 0; k < 4; ++k) rno[k] = rcy * co[k];
                    float o[4][4];
                    [[maybe_unused]] float ma[4][4];
                    float4 cc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { cc[k] = lds128(xrow + (16 + q) * X_PLANE + xo_slot[k]); cdep[k] = cc[k].w; }
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const float2 sa[4] = {SA[0][h2], SA[1][h2], SA[2][h2], SA[3][h2]};
                        const float2 sb[4] = {SB[0][h2], SB[1][h2], SB[2][h2], SB[3][h2]};
                        float2 hA[4], hb[4];
                        hsum9x2(sa, hA);
                        hsum9x2(sb, hb);
                        // mean2(A') g + mean2(b') + C = (hA g + hb) / N2 + C
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 c2 = h2 == 0 ? make_float2(cc[k].x, cc[k].y) : make_float2(cc[k].z, cc[k].w);
                            const float2 r = __ffma2_rn(__ffma2_rn(hA[k], dup2(gc[k]), hb[k]), dup2(rno[k]), c2);
                            o[k][2 * h2] = r.x; o[k][2 * h2 + 1] = r.y;
                            if constexpr (SAVE_MA) { const float2 m = __fmul2_rn(hA[k], dup2(rno[k])); ma[k][2 * h2] = m.x; ma[k][2 * h2 + 1] = m.y; }
                        }
                    }
                    if constexpr (SAVE_MA) {
                        // mean2(A'): d out_o / d guide at fixed statistics; the adjoint's direct term is sum_o gx_o * mean2(A'_o)
                        float4* mrow = oma + (unsigned)(yo * W);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (co[k] > 0.f) mrow[k] = make_float4(ma[k][0], ma[k][1], ma[k][2], ma[k][3]);
                    }
                    if constexpr (!OUT_BF) {
                        float4* orow = o32 + (unsigned)(yo * W);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (co[k] > 0.f) orow[k] = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                    } else {
                        // C8 bf16 pixel vectors: the two half-warps hold the two quads of oct w2; lanes of half 0
                        // assemble columns 0-1, lanes of half 1 columns 2-3
                        float r[2][4];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                r[i][c] = __shfl_xor_sync(0xffffffffu, h == 0 ? o[2 + i][c] : o[i][c], 16);
                        uint4* orow = o16 + (unsigned)(yo * W);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int kk = h == 0 ? i : 2 + i;
                            // (static register indices: a run-time o[kk] would put the array in local memory)
                            const float4 mine = h == 0 ? make_float4(o[i][0], o[i][1], o[i][2], o[i][3])
                                                       : make_float4(o[2 + i][0], o[2 + i][1], o[2 + i][2], o[2 + i][3]);
                            const float4 peer = make_float4(r[i][0], r[i][1], r[i][2], r[i][3]);
                            const bool ok = (4 * j + kk < OUTW) && (xo + kk < W);
                            if (ok) orow[kk] = h == 0 ? bf8_pack(mine, peer) : bf8_pack(peer, mine);
                        }
                    }
                }
                // this row of the exchange buffer has been read (A', b', C): the A' / b' loads were consumed by the TMEM stores
                // above (completed by their wait), the C loads are made operands of the arrive
                mbar_arrive_after_loads((sbase + OFF_BARS + (uint32_t)offsetof(Bars, x_empty) + 8u * (half)), rt_zero, cdep[0], cdep[1], cdep[2], cdep[3], cdep[0]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            gp0 += (uint32_t)((n1 + 1) >> 1);
        }
        GX_PROF_END(1);
    } else if (warp < 12) {
        // ============================ EP: accumulators -> (A', b', C) exchange rows ============================
        const int quarter = warp & 3, m = quarter * 32 + lane, half = m >> 6, pc = m & 63;
        const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16);
        unsigned char* xbuf = smem + OFF_X;
        const int xoff = half * X_BYTES + xslot(pc >> 2, pc & 3) * 16;
        GX_PROF_DECL;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xs = x0 - 4 + pc;
            const size_t sb = (size_t)ck.b * plane;
            const size_t bplane = (size_t)p.B * plane;
            const int n1 = rows + 8, npairs = (n1 + 1) >> 1;
            for (int pl = 0; pl < npairs; ++pl) {
                const uint32_t gpair = gp0 + (uint32_t)pl;
                const int ys = y0 - 4 + 2 * pl + half;
                float mx = 0.f, i1 = 0.f, i2 = 0.f;
                if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
                    const size_t o = sb + (size_t)ys * W + xs;
                    mx = __ldg(p.stats + o); i1 = __ldg(p.stats + bplane + o); i2 = __ldg(p.stats + 2 * bplane + o);
                }
                GX_WAIT(mbar_wait(smem_u32(&bars->d_full), gpair & 1u));
                tc_fence_after();
                float a[32], t[32], bq[32];
                tmem_ld32(tacc + 0, a);
                tmem_ld32(tacc + 32, t);
                {
                    const float2 i12 = dup2(i1), i22 = dup2(i2);
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {                                       // A' = inv1 P + inv2 Q
                        const float2 r = __ffma2_rn(i22, make_float2(t[c], t[c + 1]), __fmul2_rn(i12, make_float2(a[c], a[c + 1])));
                        a[c] = r.x; a[c + 1] = r.y;
                    }
                }
                tmem_ld32(tacc + 64, bq);
                {
                    const float2 nmx = dup2(-mx);
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {                                       // b' = R - mx A'
                        const float2 r = __ffma2_rn(nmx, make_float2(a[c], a[c + 1]), make_float2(bq[c], bq[c + 1]));
                        bq[c] = r.x; bq[c + 1] = r.y;
                    }
                }
                tmem_ld32(tacc + 96, t);
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->d_empty));
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float2 r = __fadd2_rn(make_float2(t[c], t[c + 1]), *reinterpret_cast<const float2*>(&bars->bias[c]));
                    t[c] = r.x; t[c + 1] = r.y;
                }
                GX_WAIT(mbar_wait(smem_u32(&bars->x_empty[half]), (gpair & 1u) ^ 1u));     // L2 took this row of pair gpair-1
                unsigned char* xr_ = xbuf + xoff;
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    *reinterpret_cast<float4*>(xr_ + qq * X_PLANE) = make_float4(a[qq * 4], a[qq * 4 + 1], a[qq * 4 + 2], a[qq * 4 + 3]);
                    *reinterpret_cast<float4*>(xr_ + (8 + qq) * X_PLANE) = make_float4(bq[qq * 4], bq[qq * 4 + 1], bq[qq * 4 + 2], bq[qq * 4 + 3]);
                    *reinterpret_cast<float4*>(xr_ + (16 + qq) * X_PLANE) = make_float4(t[qq * 4], t[qq * 4 + 1], t[qq * 4 + 2], t[qq * 4 + 3]);
                }
                mbar_arrive(smem_u32(&bars->x_full[half]));
            }
            gp0 += (uint32_t)npairs;
        }
        GX_PROF_END(2);
    } else if (warp == MMA_WARP) {
        // ============================ MMA issuer ============================
        const uint32_t w_base = smem_u32(smem + OFF_W), aop_base = smem_u32(smem + OFF_AOP), z_base = smem_u32(smem + OFF_Z);
        const uint32_t id64 = tc_idesc(64u, 2u), id32 = tc_idesc(32u, 2u);
        uint32_t zc0 = 0u, zc1 = 0u;                           // completed z-staging phases per buffer (scalars: no local array)
        GX_PROF_DECL;
        while (walk.next(ck)) {
            const int n1 = ck.rows + 8, npairs = (n1 + 1) >> 1;
            for (int pl = 0; pl < npairs; ++pl) {
                const uint32_t gpair = gp0 + (uint32_t)pl, buf = gpair & 1u;
                const bool zvalid = 2 * pl + 1 >= 8 && 2 * pl - 8 < ck.rows;      // any of the two output rows inside the chunk
                GX_WAIT(mbar_wait(smem_u32(&bars->aop_full[buf]), (gpair >> 1) & 1u));
                if (gpair > 0) GX_WAIT(mbar_wait(smem_u32(&bars->d_empty), (gpair - 1) & 1u));
                if (zvalid) { mbar_wait(smem_u32(&bars->z_full[buf]), (buf ? zc1 : zc0) & 1u); if (buf) ++zc1; else ++zc0; }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = aop_base + buf * AOP_BYTES;
#pragma unroll
                    for (int s = 0; s < 4; ++s)          // [P|Q] = cov [Wa|Wb]^T
                        tc_mma_tf32(tmem_base + 0, make_desc(a0 + 2 * s * AOP_PLANE, AOP_PLANE, AOP_SBO),
                                    make_desc(w_base + s * 2048, 1024, 128), id64, s ? 1u : 0u);
#pragma unroll
                    for (int s = 0; s < 4; ++s)          // R = mean_z (Wa+Wb)^T
                        tc_mma_tf32(tmem_base + 64, make_desc(a0 + (8 + 2 * s) * AOP_PLANE, AOP_PLANE, AOP_SBO),
                                    make_desc(w_base + 8192 + s * 1024, 512, 128), id32, s ? 1u : 0u);
                    if (zvalid) {
#pragma unroll
                        for (int s = 0; s < 4; ++s)      // C = z[output rows] Wc^T
                            tc_mma_tf32(tmem_base + 96, make_desc(z_base + buf * Z_BYTES + 2 * s * 2048, 2048, 128),
                                        make_desc(w_base + 12288 + s * 1024, 512, 128), id32, s ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&bars->d_full));
                    tc_commit(smem_u32(&bars->aop_empty[buf]));
                }
                __syncwarp();
            }
            gp0 += (uint32_t)npairs;
        }
        GX_PROF_END(3);
    } else if (warp == MMA_WARP + 1) {
        // ============================ producer (one elected thread issues the bulk copies) ============================
        //  * raw-row ring: the entering row (8 quad planes + guide, columns x0-8 .. x0+55 clipped to the image) of every L1
        //    iteration, RS stages ahead;
        //  * z staging: the raw rows of the two OUTPUT rows of each level-1 pair (A operand of the Wc GEMM), columns x0..x0+63.
        uint32_t rcount = 0;
        GX_PROF_DECL;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xlo = max(0, x0 - 8), xhi = min(W, x0 + 56);              // raw columns inside the image
            const uint32_t ring_px = (uint32_t)(xhi - xlo), ring_off = (uint32_t)(xlo - (x0 - 8));
            const int npx = min(64, W - x0);
            const int n1 = rows + 8, npairs = (n1 + 1) >> 1, nt = rows + 16;
            const float4* fsrc = reinterpret_cast<const float4*>(p.feat) + (size_t)ck.b * 8 * plane;
            const float* gsrc = p.guide + (size_t)ck.b * plane;
            // z staging of pair pl is issued 12 + 2 pl ring rows into the chunk: by then the MMAs of pair pl-2 (whose
            // completion frees the stage) have normally retired, so the wait below does not hold up the ring
            for (int t = 0; t < nt + 4 || ((t - 12) >> 1) < npairs; ++t) {
                if (t < nt) {
                    const int yr = y0 - 8 + t;
                    const uint32_t st = rcount % RS;
                    GX_WAIT(mbar_wait(smem_u32(&bars->r_empty[st]), ((rcount / RS) & 1u) ^ 1u));
                    if (t < RS && ring_px < 64u) {
                        // first use of this stage in the chunk: the columns the copies of this chunk never write (outside
                        // the image) may hold rows of another strip — clear them, so that L1 needs no column masks.  The
                        // stores are ordered before L1's reads by the __syncwarp + the elected lane's arrive on r_full.
                        unsigned char* dstz = smem + OFF_R + st * R_BYTES;
                        const uint32_t nstale = 64u - ring_px;                     // columns [0, ring_off) and [ring_off + ring_px, 64)
                        for (uint32_t i = lane; i < nstale * 9u; i += 32u) {
                            const uint32_t pl9 = i / nstale, ci = i - pl9 * nstale;
                            const uint32_t col = ci < ring_off ? ci : ci + ring_px;
                            if (pl9 < 8u) *reinterpret_cast<float4*>(dstz + pl9 * 1024 + col * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
                            else *reinterpret_cast<float*>(dstz + 8192 + col * 4) = 0.f;
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        const uint32_t bar = smem_u32(&bars->r_full[st]);
                        if (yr >= 0 && yr < H) {
                            mbar_expect_tx(bar, ring_px * (8 * 16 + 4));
                            const uint32_t dst = smem_u32(smem + OFF_R) + st * R_BYTES;
#pragma unroll
                            for (int qq = 0; qq < 8; ++qq)
                                bulk_g2s(dst + qq * 1024 + ring_off * 16, fsrc + qq * plane + (size_t)yr * W + xlo, ring_px * 16, bar);
                            bulk_g2s(dst + 8192 + ring_off * 4, gsrc + (size_t)yr * W + xlo, ring_px * 4, bar);
                        } else {
                            mbar_arrive(bar);                                    // a row outside the image: no data, the phase still turns
                        }
                    }
                    __syncwarp();
                    ++rcount;
                }
                const int pl = (t - 12) >> 1;
                if (t >= 12 && ((t - 12) & 1) == 0 && pl < npairs) {
                    const uint32_t gpair = gp0 + (uint32_t)pl, buf = gpair & 1u;
                    const int yoA = y0 + 2 * pl - 8, yoB = yoA + 1;
                    const bool vA = yoA >= y0 && yoA < y0 + rows, vB = yoB >= y0 && yoB < y0 + rows;
                    // The stage was last read by the MMAs of pair gpair-2, whose completion is aop_empty[buf]'s previous
                    // phase.  Waited for on EVERY pair (also those without an output row): a parity wait can only tell
                    // adjacent phases apart, so the producer must never run more than one phase ahead of the issuer.
                    GX_WAIT(mbar_wait(smem_u32(&bars->aop_empty[buf]), ((gpair >> 1) & 1u) ^ 1u));
                    if ((vA || vB) && elect_one()) {
                        const uint32_t bar = smem_u32(&bars->z_full[buf]);
                        const uint32_t row_bytes = (uint32_t)npx * 16;
                        mbar_expect_tx(bar, row_bytes * 8 * ((vA ? 1 : 0) + (vB ? 1 : 0)));
                        const uint32_t dst = smem_u32(smem + OFF_Z) + buf * Z_BYTES;
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            if (vA) bulk_g2s(dst + qq * 2048, fsrc + qq * plane + (size_t)yoA * W + x0, row_bytes, bar);
                            if (vB) bulk_g2s(dst + qq * 2048 + 1024, fsrc + qq * plane + (size_t)yoB * W + x0, row_bytes, bar);
                        }
                    }
                    __syncwarp();
                }
            }
            gp0 += (uint32_t)npairs;
        }
        GX_PROF_END(4);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace gx
}  // namespace paif

using namespace paif;

#ifdef PAIF_TC_PROFILE
extern "C" int paif_debug_gx_counters(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, gx::gx_prof, sizeof(gx::gx_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(gx::gx_prof, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int paif_gf_mix_supported(int C, int H, int W) { return C == 32 && H > 9 && W > 9 && W % 4 == 0; }

static int gf_mix_launch(const float* feat, const float* residue, const float* stats, const void* wpack,
                         const float* bias, void* out, int out_bf16, float* mean_a, int C, int B, int H, int W, void* stream,
                         const char* what) {
    PAIF_REQUIRE(feat && residue && stats && wpack && out, "null pointer");
    PAIF_REQUIRE(B > 0 && H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    if (!paif_gf_mix_supported(C, H, W)) { set_error("%s: needs C = 32 and W %% 4 == 0", what); return PAIF_ENOTSUP; }
    PAIF_REQUIRE(((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(residue) | reinterpret_cast<uintptr_t>(stats) |
                   reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(mean_a)) & 15) == 0,
                 "pointers must be 16-byte aligned");
    static unsigned long long attr_done = 0;
    int dev;
    if (attr_needed(attr_done, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(gx::gf_mix_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gx::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gx::gf_mix_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gx::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gx::gf_mix_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gx::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gx::gf_mix_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gx::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gf_mix smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_mark(attr_done, dev);
    }
    gx::Params p;
    p.feat = feat; p.guide = residue; p.stats = stats; p.wpack = wpack; p.bias = bias; p.out = out; p.mean_a = mean_a;
    p.B = B; p.H = H; p.W = W; p.nstrips = cdiv(W, gx::OUTW);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Row chunks (the running sums restart per chunk, which also bounds their rounding drift): ~160 rows when that
    // gives every SM work (16 + 8 halo rows of recompute per chunk); down to 32 rows when a small batch would
    // otherwise leave SMs idle (one 480x640 frame: 14 strips x 10 chunks of 48 rows).  The chunk grid is a function of
    // (B * strips, H) only, so equal-shaped launches are bit-identical whatever the batch position of an image.
    const long long strips = (long long)B * p.nstrips;
    int n = cdiv(H, 160);
    if (strips * n < 2LL * sms) {
        long long want = (2LL * sms + strips - 1) / strips;
        const int most = cdiv(H, 32);
        n = (int)(want < most ? want : most);
    }
    p.RC = cdiv(H, n);
    p.nchunks = cdiv(H, p.RC);
    PAIF_REQUIRE(strips * p.nchunks < (1ll << 30), "problem too large");
    p.nitems = (int)(strips * p.nchunks);
    const int grid = p.nitems < sms ? p.nitems : sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (mean_a) {
        if (out_bf16) gx::gf_mix_kernel<true, true><<<grid, gx::NT, gx::SMEM_BYTES, st>>>(p);
        else gx::gf_mix_kernel<false, true><<<grid, gx::NT, gx::SMEM_BYTES, st>>>(p);
    } else {
        if (out_bf16) gx::gf_mix_kernel<true, false><<<grid, gx::NT, gx::SMEM_BYTES, st>>>(p);
        else gx::gf_mix_kernel<false, false><<<grid, gx::NT, gx::SMEM_BYTES, st>>>(p);
    }
    return check_launch(what);
}

extern "C" int paif_gf_mix_forward(const float* feat, const float* residue, const float* stats, const void* wpack,
                                   const float* bias, void* out, int out_bf16, int C, int B, int H, int W, void* stream) {
    return gf_mix_launch(feat, residue, stats, wpack, bias, out, out_bf16, nullptr, C, B, H, W, stream, "paif_gf_mix_forward");
}

extern "C" int paif_gf_mix_forward_save(const float* feat, const float* residue, const float* stats, const void* wpack,
                                        const float* bias, void* out, int out_bf16, float* mean_a,
                                        int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(mean_a, "null pointer");
    return gf_mix_launch(feat, residue, stats, wpack, bias, out, out_bf16, mean_a, C, B, H, W, stream, "paif_gf_mix_forward_save");
}
