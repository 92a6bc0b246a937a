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
#include "tc_ptx.cuh"

namespace paif {
namespace gx {

constexpr int OUTW = 48;                       // output columns per strip (64 raw -> 56 level-1 -> 48 output)
constexpr int NWARPS = 14, NT = NWARPS * 32;
constexpr int MMA_WARP = 12;            // warp 13: producer
constexpr int AOP_SBO = 144;                   // core-matrix pitch of the A operand (128 B + 16 B pad: conflict-free 4-column stores)
constexpr int AOP_PLANE = 16 * AOP_SBO;        // one 16-byte K chunk (4 channels) x 128 rows
constexpr int AOP_BYTES = 16 * AOP_PLANE;      // planes 0-7: cov quads, 8-15: mean_z quads
constexpr int X_PLANE = 2048;                  // exchange: [plane][row half][64 slots][16 B]
constexpr int X_BYTES = 24 * X_PLANE;          // planes 0-7 A', 8-15 b', 16-23 C
constexpr int Z_BYTES = 8 * 2048;              // 8 quad planes x 128 rows x 16 B (rows 0-63: first output row, 64-127: second)
constexpr int W_BYTES = 16384;                 // [Wa|Wb] 8 KB, Wa+Wb 4 KB, Wc 4 KB (TF32, UMMA B tiles)
constexpr int OFF_W = 0, OFF_AOP = OFF_W + W_BYTES, OFF_X = OFF_AOP + 2 * AOP_BYTES, OFF_Z = OFF_X + 2 * X_BYTES,
              OFF_BARS = OFF_Z + 2 * Z_BYTES, SMEM_BYTES = OFF_BARS + 1024;
constexpr int RING_COL0 = 128, RING_SLOTS = 9;

struct Bars {
    uint64_t aop_full[2], aop_empty[2], z_full[2], x_full[2], x_empty[2], d_full, d_empty;
    uint32_t tmem_base;
    alignas(16) float bias[32];
};
static_assert(sizeof(Bars) <= 1024, "barrier block");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
    const float* feat; const float* guide; const float* stats; const void* wpack; const float* bias;
    void* out;
    int B, H, W, nstrips;
    int RC, nchunks;           // rows per chunk / chunks per strip: a function of the shape only (see the launcher)
    int nitems;                // B * nstrips * nchunks work items, split into contiguous ranges over the CTAs
};

__device__ __forceinline__ float win_count(int p, int n) {
    const int lo = p - 4 < 0 ? 0 : p - 4, hi = p + 4 > n - 1 ? n - 1 : p + 4;
    return (float)(hi - lo + 1);
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
__device__ __forceinline__ void hsum9c(const float (&s)[4][4], int c, float (&o)[4]) {
    const float a[4] = {s[0][c], s[1][c], s[2][c], s[3][c]};
    hsum9(a, o);
}
// 4 consecutive floats of a plane row (x multiple of 4, W multiple of 4): zero outside [0, W)
__device__ __forceinline__ void ld_cols4(const float* __restrict__ row, int x, int W, float (&v)[4]) {
    if (x >= 0 && x < W) { const float4 t = __ldg(reinterpret_cast<const float4*>(row + x)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else { v[0] = v[1] = v[2] = v[3] = 0.f; }
}
// the 4 channels of one quad at pixels x..x+3 of a C4 map row (`row` -> (quad plane, row y, pixel 0)); zero outside
__device__ __forceinline__ void ld_quad(const float* __restrict__ row, int x, int W, float (&z)[4][4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        if (xx >= 0 && xx < W) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(row + (size_t)xx * 4));
            z[k][0] = t.x; z[k][1] = t.y; z[k][2] = t.z; z[k][3] = t.w;
        } else { z[k][0] = z[k][1] = z[k][2] = z[k][3] = 0.f; }
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

template <bool OUT_BF>
__global__ void __maxnreg__(144)          // 448 threads x 144 registers: one CTA per SM
gf_mix_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BARS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->aop_full[i]), 128);
            mbar_init(smem_u32(&bars->aop_empty[i]), 1);
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->x_full[i]), 128);
            mbar_init(smem_u32(&bars->x_empty[i]), 128);
        }
        mbar_init(smem_u32(&bars->d_full), 1);
        mbar_init(smem_u32(&bars->d_empty), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) bars->bias[tid] = p.bias ? p.bias[tid] : 0.f;
    for (int i = tid; i < W_BYTES / 16; i += NT)
        reinterpret_cast<uint4*>(smem + OFF_W)[i] = __ldg(reinterpret_cast<const uint4*>(p.wpack) + i);
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
        unsigned char* aop = smem + OFF_AOP;
        // byte offset of row m = half*64 + 4j + k inside a plane: (m >> 3) * SBO + (m & 7) * 16
        const int aoff = (j >> 1) * AOP_SBO + (j & 1) * 64;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xr = x0 - 8 + 4 * j, xs = xr + 4;
            const float* zp = p.feat + ((size_t)ck.b * 8 + q) * plane * 4;
            const float* gp = p.guide + (size_t)ck.b * plane;
            const float* mxp = p.stats + (size_t)ck.b * plane;
            float cs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                cs[k] = (xs + k >= 0 && xs + k < W && 4 * j + k < 56) ? __frcp_rn(win_count(xs + k, W)) : 0.f;
            float Sz[4][4], Sgz[4][4], zn[4][4], zq[4][4], gn[4], gq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                gq[k] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) { Sz[k][c] = Sgz[k][c] = zq[k][c] = 0.f; }
            }
            {
                const int yr = y0 - 8;
                if (yr >= 0) { ld_quad(zp + (size_t)yr * W * 4, xr, W, zn); ld_cols4(gp + (size_t)yr * W, xr, W, gn); }
                else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { gn[k] = 0.f; zn[k][0] = zn[k][1] = zn[k][2] = zn[k][3] = 0.f; }
                }
            }
            const int n1 = rows + 8, nt = rows + 16;
            for (int t = 0; t < nt; ++t) {
                const int yr = y0 - 8 + t;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        Sz[k][c] += zn[k][c] - zq[k][c];
                        Sgz[k][c] = __fmaf_rn(-gq[k], zq[k][c], __fmaf_rn(gn[k], zn[k][c], Sgz[k][c]));
                    }
                {   // rows of iteration t+1: entering yr+1, leaving yr-8
                    const int yn = yr + 1, yl = yr - 8;
                    if (t + 1 < nt && yn >= 0 && yn < H) { ld_quad(zp + (size_t)yn * W * 4, xr, W, zn); ld_cols4(gp + (size_t)yn * W, xr, W, gn); }
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { gn[k] = 0.f; zn[k][0] = zn[k][1] = zn[k][2] = zn[k][3] = 0.f; }
                    }
                    if (t + 1 >= 9 && yl >= 0 && yl < H) { ld_quad(zp + (size_t)yl * W * 4, xr, W, zq); ld_cols4(gp + (size_t)yl * W, xr, W, gq); }
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { gq[k] = 0.f; zq[k][0] = zq[k][1] = zq[k][2] = zq[k][3] = 0.f; }
                    }
                }
                if (t < 8) continue;
                const int r1 = t - 8, ys = yr - 4, half = r1 & 1;
                const uint32_t gpair = gp0 + (uint32_t)(r1 >> 1), buf = gpair & 1u;
                if (half == 0) mbar_wait(smem_u32(&bars->aop_empty[buf]), ((gpair >> 1) & 1u) ^ 1u);
                float mx[4], rn[4];
                if (ys >= 0 && ys < H) {
                    ld_cols4(mxp + (size_t)ys * W, xs, W, mx);
                    const float rcy = __frcp_rn(win_count(ys, H));
#pragma unroll
                    for (int k = 0; k < 4; ++k) rn[k] = rcy * cs[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mx[k] = rn[k] = 0.f;
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
                        cov[k][c] = __fmaf_rn(-mx[k], mz[k][c], bg[k] * rn[k]);
                    }
                }
                unsigned char* base = aop + buf * AOP_BYTES + half * (8 * AOP_SBO) + aoff;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    *reinterpret_cast<float4*>(base + q * AOP_PLANE + k * 16) = make_float4(cov[k][0], cov[k][1], cov[k][2], cov[k][3]);
                    *reinterpret_cast<float4*>(base + (8 + q) * AOP_PLANE + k * 16) = make_float4(mz[k][0], mz[k][1], mz[k][2], mz[k][3]);
                }
                if (half == 1 || r1 == n1 - 1) {
                    fence_proxy_async();
                    mbar_arrive(smem_u32(&bars->aop_full[buf]));
                }
            }
            gp0 += (uint32_t)((n1 + 1) >> 1);
        }
    } else if (warp < 8) {
        // ============================ L2: level-2 box filters -> output ============================
        const int w2 = warp & 3, h = lane >> 4, j = lane & 15, q = 2 * w2 + h;
        const unsigned char* xbuf = smem + OFF_X;
        const uint32_t ring = tmem_base + ((uint32_t)(w2 * 32) << 16) + RING_COL0;
        int xo_slot[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xo_slot[k] = xslot(j, k) * 16;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xo = x0 + 4 * j;
            const float* gp = p.guide + (size_t)ck.b * plane;
            float co[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                co[k] = (4 * j + k < OUTW && xo + k < W) ? __frcp_rn(win_count(xo + k, W)) : 0.f;
            float S[32];                                           // [A' | b'] x [column k][channel c]: index e*16 + k*4 + c
#pragma unroll
            for (int i = 0; i < 32; ++i) S[i] = 0.f;
            const int n1 = rows + 8, npairs = (n1 + 1) >> 1;
            int slot = 0;
            for (int pl = 0; pl < npairs; ++pl) {
                const uint32_t gpair = gp0 + (uint32_t)pl, xb = gpair & 1u;
                mbar_wait(smem_u32(&bars->x_full[xb]), (gpair >> 1) & 1u);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r1 = 2 * pl + half;
                    if (r1 < n1) {
                        const unsigned char* xrow = xbuf + xb * X_BYTES + half * 1024;
                        float nw[32];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 a = *reinterpret_cast<const float4*>(xrow + q * X_PLANE + xo_slot[k]);
                            const float4 bb = *reinterpret_cast<const float4*>(xrow + (8 + q) * X_PLANE + xo_slot[k]);
                            nw[k * 4 + 0] = a.x; nw[k * 4 + 1] = a.y; nw[k * 4 + 2] = a.z; nw[k * 4 + 3] = a.w;
                            nw[16 + k * 4 + 0] = bb.x; nw[16 + k * 4 + 1] = bb.y; nw[16 + k * 4 + 2] = bb.z; nw[16 + k * 4 + 3] = bb.w;
                        }
                        if (r1 >= RING_SLOTS) {
                            float od[32];
                            tmem_ld32(ring + slot * 32, od);
#pragma unroll
                            for (int i = 0; i < 32; ++i) S[i] += nw[i] - od[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) S[i] += nw[i];
                        }
                        tmem_st32(ring + slot * 32, nw);
                        slot = slot == RING_SLOTS - 1 ? 0 : slot + 1;
                        if (r1 >= 8) {
                            const int yo = y0 + r1 - 8;
                            float g[4], rno[4];
                            ld_cols4(gp + (size_t)yo * W, xo, W, g);
                            const float rcy = __frcp_rn(win_count(yo, H));
#pragma unroll
                            for (int k = 0; k < 4; ++k) rno[k] = rcy * co[k];
                            float o[4][4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float sa[4] = {S[c], S[4 + c], S[8 + c], S[12 + c]};
                                const float sb[4] = {S[16 + c], S[20 + c], S[24 + c], S[28 + c]};
                                float hA[4], hb[4];
                                hsum9(sa, hA);
                                hsum9(sb, hb);
#pragma unroll
                                for (int k = 0; k < 4; ++k) o[k][c] = __fmaf_rn(hA[k] * rno[k], g[k], hb[k] * rno[k]);
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float4 cc = *reinterpret_cast<const float4*>(xrow + (16 + q) * X_PLANE + xo_slot[k]);
                                o[k][0] += cc.x; o[k][1] += cc.y; o[k][2] += cc.z; o[k][3] += cc.w;
                            }
                            if constexpr (!OUT_BF) {
                                float* orow = static_cast<float*>(p.out) + (((size_t)ck.b * 8 + q) * plane + (size_t)yo * W) * 4;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (co[k] > 0.f)
                                        *reinterpret_cast<float4*>(orow + (size_t)(xo + k) * 4) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                            } else {
                                // C8 bf16 pixel vectors: the two half-warps hold the two quads of oct w2; lanes of half 0
                                // assemble columns 0-1, lanes of half 1 columns 2-3
                                float r[2][4];
#pragma unroll
                                for (int i = 0; i < 2; ++i)
#pragma unroll
                                    for (int c = 0; c < 4; ++c)
                                        r[i][c] = __shfl_xor_sync(0xffffffffu, h == 0 ? o[2 + i][c] : o[i][c], 16);
                                uint4* orow = static_cast<uint4*>(p.out) + ((size_t)ck.b * 4 + w2) * plane + (size_t)yo * W;
#pragma unroll
                                for (int i = 0; i < 2; ++i) {
                                    const int kk = h == 0 ? i : 2 + i;
                                    const float4 mine = make_float4(o[kk][0], o[kk][1], o[kk][2], o[kk][3]);
                                    const float4 peer = make_float4(r[i][0], r[i][1], r[i][2], r[i][3]);
                                    const bool ok = (4 * j + kk < OUTW) && (xo + kk < W);
                                    if (ok) orow[xo + kk] = h == 0 ? bf8_pack(mine, peer) : bf8_pack(peer, mine);
                                }
                            }
                        }
                    }
                }
                mbar_arrive(smem_u32(&bars->x_empty[xb]));
            }
            gp0 += (uint32_t)npairs;
        }
    } else if (warp < 12) {
        // ============================ EP: accumulators -> (A', b', C) exchange rows ============================
        const int quarter = warp & 3, m = quarter * 32 + lane, half = m >> 6, pc = m & 63;
        const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16);
        unsigned char* xbuf = smem + OFF_X;
        const int xoff = half * 1024 + xslot(pc >> 2, pc & 3) * 16;
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int xs = x0 - 4 + pc;
            const size_t sb = (size_t)ck.b * plane;
            const size_t bplane = (size_t)p.B * plane;
            const int n1 = rows + 8, npairs = (n1 + 1) >> 1;
            for (int pl = 0; pl < npairs; ++pl) {
                const uint32_t gpair = gp0 + (uint32_t)pl, xb = gpair & 1u;
                const int ys = y0 - 4 + 2 * pl + half;
                float mx = 0.f, i1 = 0.f, i2 = 0.f;
                if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
                    const size_t o = sb + (size_t)ys * W + xs;
                    mx = __ldg(p.stats + o); i1 = __ldg(p.stats + bplane + o); i2 = __ldg(p.stats + 2 * bplane + o);
                }
                mbar_wait(smem_u32(&bars->d_full), gpair & 1u);
                tc_fence_after();
                float a[32], t[32], bq[32];
                tmem_ld32(tacc + 0, a);
                tmem_ld32(tacc + 32, t);
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] = __fmaf_rn(i2, t[c], i1 * a[c]);          // A' = inv1 P + inv2 Q
                tmem_ld32(tacc + 64, bq);
#pragma unroll
                for (int c = 0; c < 32; ++c) bq[c] = __fmaf_rn(-mx, a[c], bq[c]);            // b' = R - mx A'
                tmem_ld32(tacc + 96, t);
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->d_empty));
#pragma unroll
                for (int c = 0; c < 32; ++c) t[c] += bars->bias[c];
                mbar_wait(smem_u32(&bars->x_empty[xb]), ((gpair >> 1) & 1u) ^ 1u);
                unsigned char* xr_ = xbuf + xb * X_BYTES + xoff;
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    *reinterpret_cast<float4*>(xr_ + qq * X_PLANE) = make_float4(a[qq * 4], a[qq * 4 + 1], a[qq * 4 + 2], a[qq * 4 + 3]);
                    *reinterpret_cast<float4*>(xr_ + (8 + qq) * X_PLANE) = make_float4(bq[qq * 4], bq[qq * 4 + 1], bq[qq * 4 + 2], bq[qq * 4 + 3]);
                    *reinterpret_cast<float4*>(xr_ + (16 + qq) * X_PLANE) = make_float4(t[qq * 4], t[qq * 4 + 1], t[qq * 4 + 2], t[qq * 4 + 3]);
                }
                mbar_arrive(smem_u32(&bars->x_full[xb]));
            }
            gp0 += (uint32_t)npairs;
        }
    } else if (warp == MMA_WARP) {
        // ============================ MMA issuer ============================
        const uint32_t w_base = smem_u32(smem + OFF_W), aop_base = smem_u32(smem + OFF_AOP), z_base = smem_u32(smem + OFF_Z);
        const uint32_t id64 = tc_idesc(64u, 2u), id32 = tc_idesc(32u, 2u);
        uint32_t zcnt[2] = {0u, 0u};
        while (walk.next(ck)) {
            const int n1 = ck.rows + 8, npairs = (n1 + 1) >> 1;
            for (int pl = 0; pl < npairs; ++pl) {
                const uint32_t gpair = gp0 + (uint32_t)pl, buf = gpair & 1u;
                const bool zvalid = 2 * pl + 1 >= 8 && 2 * pl - 8 < ck.rows;      // any of the two output rows inside the chunk
                mbar_wait(smem_u32(&bars->aop_full[buf]), (gpair >> 1) & 1u);
                if (gpair > 0) mbar_wait(smem_u32(&bars->d_empty), (gpair - 1) & 1u);
                if (zvalid) { mbar_wait(smem_u32(&bars->z_full[buf]), zcnt[buf] & 1u); ++zcnt[buf]; }
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
    } else {
        // ============================ producer: raw z rows of the output rows ============================
        while (walk.next(ck)) {
            const int x0 = ck.strip * OUTW, y0 = ck.y0, rows = ck.rows;
            const int npx = min(64, W - x0);
            const int n1 = rows + 8, npairs = (n1 + 1) >> 1;
            for (int pl = 0; pl < npairs; ++pl) {
                const uint32_t gpair = gp0 + (uint32_t)pl, buf = gpair & 1u;
                const int yoA = y0 + 2 * pl - 8, yoB = yoA + 1;
                const bool vA = yoA >= y0 && yoA < y0 + rows, vB = yoB >= y0 && yoB < y0 + rows;
                if (!(vA || vB)) continue;
                // the stage was last read by the MMAs of pair gpair-2, whose completion is aop_empty[buf]'s previous phase
                mbar_wait(smem_u32(&bars->aop_empty[buf]), ((gpair >> 1) & 1u) ^ 1u);
                if (elect_one()) {
                    const uint32_t bar = smem_u32(&bars->z_full[buf]);
                    const uint32_t row_bytes = (uint32_t)npx * 16;
                    mbar_expect_tx(bar, row_bytes * 8 * ((vA ? 1 : 0) + (vB ? 1 : 0)));
                    const uint32_t dst = smem_u32(smem + OFF_Z) + buf * Z_BYTES;
                    const float4* src = reinterpret_cast<const float4*>(p.feat) + (size_t)ck.b * 8 * plane + x0;
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        if (vA) bulk_g2s(dst + qq * 2048, src + qq * plane + (size_t)yoA * W, row_bytes, bar);
                        if (vB) bulk_g2s(dst + qq * 2048 + 1024, src + qq * plane + (size_t)yoB * W, row_bytes, bar);
                    }
                }
                __syncwarp();
            }
            gp0 += (uint32_t)((n1 + 1) >> 1);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace gx
}  // namespace paif

using namespace paif;

extern "C" int paif_gf_mix_supported(int C, int H, int W) { return C == 32 && H > 9 && W > 9 && W % 4 == 0; }

extern "C" int paif_gf_mix_forward(const float* feat, const float* residue, const float* stats, const void* wpack,
                                   const float* bias, void* out, int out_bf16, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && residue && stats && wpack && out, "null pointer");
    PAIF_REQUIRE(B > 0 && H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    if (!paif_gf_mix_supported(C, H, W)) { set_error("paif_gf_mix_forward: needs C = 32 and W %% 4 == 0"); return PAIF_ENOTSUP; }
    PAIF_REQUIRE(((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(residue) | reinterpret_cast<uintptr_t>(stats) |
                   reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "pointers must be 16-byte aligned");
    static unsigned long long attr_done = 0;
    int dev;
    if (attr_needed(attr_done, &dev)) {
        cudaError_t e = cudaFuncSetAttribute(gx::gf_mix_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gx::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gx::gf_mix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gx::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gf_mix smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_mark(attr_done, dev);
    }
    gx::Params p;
    p.feat = feat; p.guide = residue; p.stats = stats; p.wpack = wpack; p.bias = bias; p.out = out;
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
    int grid = p.nitems < sms ? p.nitems : sms;
    if (out_bf16) gx::gf_mix_kernel<true><<<grid, gx::NT, gx::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    else gx::gf_mix_kernel<false><<<grid, gx::NT, gx::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    return check_launch("paif_gf_mix_forward");
}
