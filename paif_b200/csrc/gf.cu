// Guided-filter decomposition (Cell_Decom.decomposition, core/model_fusion_auto.py:522-535, with
// the un-vendored guided_filter_pytorch GuidedFilter(r=4, eps)) and its adjoint.
//
//   N = box(1); mx = box(g)/N; var = box(g*g)/N - mx^2
//   per channel c, eps in {1e-3,1e-4}:  mz = box(z_c)/N; cov = box(g z_c)/N - mx mz
//       A = cov/(var+eps); b = mz - A mx;  LF_eps,c = box(A)/N * g + box(b)/N
//   box = 9x9 window sum clipped to the image.
//
// Forward and adjoint are row-marching, register-resident warp kernels (see below); all statistics stay fp32.
#include "common.cuh"

namespace paif {

constexpr float GF_EPS1 = 0.001f, GF_EPS2 = 0.0001f;

// clipped-window pixel count along one axis
__device__ __forceinline__ float win_count(int p, int n) {
    const int lo = p - 4 < 0 ? 0 : p - 4, hi = p + 4 > n - 1 ? n - 1 : p + 4;
    return (float)(hi - lo + 1);
}

// ------------------------------------------------------------------------------------------
// forward, pass 0: guide statistics shared by all 32 channels: mean_g and 1/(var_g + eps) for
// both eps.  stats = [3][B][H][W] (mean, inv1, inv2).  One 32x32 tile per CTA, halo 4.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gf_guide_stats_kernel(const float* __restrict__ guide, float* __restrict__ stats, int B, int H, int W) {
    __shared__ float sg[40][41];
    __shared__ float h1[40][33], h2[40][33];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, b = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* gp = guide + (size_t)b * plane;
    for (int i = threadIdx.x; i < 1600; i += 256) {
        const int r = i / 40, c = i - r * 40;
        const int y = y0 - 4 + r, x = x0 - 4 + c;
        sg[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? gp[(size_t)y * W + x] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1280; i += 256) {
        const int r = i >> 5, c = i & 31;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const float v = sg[r][c + k]; s1 += v; s2 += v * v; }
        h1[r][c] = s1; h2[r][c] = s2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256) {
        const int r = i >> 5, c = i & 31;
        const int y = y0 + r, x = x0 + c;
        if (y < H && x < W) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) { s1 += h1[r + k][c]; s2 += h2[r + k][c]; }
            const float n = win_count(y, H) * win_count(x, W);
            const float m = __fdiv_rn(s1, n);
            const float var = __fdiv_rn(s2, n) - m * m;
            const size_t o = (size_t)b * plane + (size_t)y * W + x;
            stats[o] = m;
            stats[(size_t)B * plane + o] = __fdiv_rn(1.f, var + GF_EPS1);
            stats[(size_t)2 * B * plane + o] = __fdiv_rn(1.f, var + GF_EPS2);
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward, pass 1: row-marching, register-resident guided filter.
//
// One WARP = one work item: (image b, channel group, 112-column strip, row chunk); a channel group
// is GF_NCH adjacent channels (a whole C4 quad).  Lane j owns 4 adjacent columns x GF_NCH
// channels.  Both box-filter levels are "vertical running sum in registers, then horizontal 9-sum
// by warp shuffles": level 1 runs on the raw rows (z, g*z), level 2 on A_e = cov/(var+eps_e),
// b_e = mean_z - A_e mean_g.  The row leaving a vertical window is re-derived (level 1: re-read
// from global/L2; level 2: recomputed from (mean_z, cov) kept in a 9-row shared-memory ring) and
// subtracted.  The sums restart at every row chunk (16 halo rows), which bounds the running-sum
// rounding drift; there is no __syncthreads anywhere.
// Each horizontal pass shifts the columns a lane owns by +4: raw columns x0-8+4j.., level-1
// columns x0-4+4j.. (valid for 4j+k <= 119), output columns x0+4j.. (valid for 4j+k <= 111).
// GF_NCH = 2 (half a quad per lane: half the registers and ring, twice the resident warps) was measured 30 %
// slower (forward 1.55 vs 1.20 ms): the per-lane costs shared by all channels (guide / statistics loads, window
// counts, addressing) double per pixel-channel and outweigh the occupancy gain.
// ------------------------------------------------------------------------------------------
constexpr int GF_NCH = 4;                  // channels per lane (2 was measured: 30 % slower)
constexpr int GF_SUBS = 4 / GF_NCH;        // channel groups per C4 quad
constexpr int GM_OUTW = 112;               // output columns per warp (two horizontal passes)
constexpr int GA_OUTW = 120;               // output columns per warp (one horizontal pass)
constexpr int GM_WPC = 2;                  // warps (work items) per CTA
constexpr int GM_RING_F4 = 9 * 2 * GF_NCH * 32;     // float4 per warp: 9 rows x (mean_z, cov) x 4 columns x GF_NCH
constexpr int GM_SMEM = GM_WPC * GM_RING_F4 * 16;

// o[k] = sum of columns (4j+k) .. (4j+k+8) of the per-lane column quadruples a[0..3]
__device__ __forceinline__ void hsum9(const float (&a)[4], float (&o)[4]) {
    const float p01 = a[0] + a[1], p23 = a[2] + a[3], p012 = p01 + a[2], full = p01 + p23;
    const float n1 = __shfl_down_sync(0xffffffffu, full, 1);      // columns +4..+7
    const float v8 = __shfl_down_sync(0xffffffffu, a[0], 2);      // column  +8
    const float v89 = __shfl_down_sync(0xffffffffu, p01, 2);      // columns +8..+9
    const float v8a = __shfl_down_sync(0xffffffffu, p012, 2);     // columns +8..+10
    const float v8b = __shfl_down_sync(0xffffffffu, full, 2);     // columns +8..+11
    o[0] = full + n1 + v8;
    o[1] = (a[1] + p23) + n1 + v89;
    o[2] = p23 + n1 + v8a;
    o[3] = a[3] + n1 + v8b;
}
// same for column c of a [4][GF_NCH] array
__device__ __forceinline__ void hsum9c(const float (&s)[4][GF_NCH], int c, float (&o)[4]) {
    const float a[4] = {s[0][c], s[1][c], s[2][c], s[3][c]};
    hsum9(a, o);
}

template <bool VEC>
__device__ __forceinline__ void ld_cols4(const float* __restrict__ row, int x, int W, float (&v)[4]) {
    if (VEC) {
        if (x >= 0 && x < W) { const float4 t = __ldg(reinterpret_cast<const float4*>(row + x)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { v[0] = v[1] = v[2] = v[3] = 0.f; }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (x + k >= 0 && x + k < W) ? __ldg(row + x + k) : 0.f;
    }
}
template <bool VEC>
__device__ __forceinline__ void st_cols4(float* __restrict__ row, int x, const float (&v)[4], const bool (&ok)[4]) {
    if (VEC && ok[0] && ok[3]) { *reinterpret_cast<float4*>(row + x) = make_float4(v[0], v[1], v[2], v[3]); return; }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (ok[k]) row[x + k] = v[k];
}

// the GF_NCH channels of this lane's group at pixels x..x+3 of one row of a C4 map; `row` points at
// (quad plane, row y, pixel 0, first channel of the group); zero outside [0, W)
__device__ __forceinline__ void ld_grp(const float* __restrict__ row, int x, int W, float (&z)[4][GF_NCH]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        if (xx >= 0 && xx < W) {
            if (GF_NCH == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(row + (size_t)xx * 4)); z[k][0] = t.x; z[k][GF_NCH - 1] = t.y; }
            else { const float4 t = __ldg(reinterpret_cast<const float4*>(row + (size_t)xx * 4)); z[k][0] = t.x; z[k][1 % GF_NCH] = t.y; z[k][2 % GF_NCH] = t.z; z[k][3 % GF_NCH] = t.w; }
        } else {
#pragma unroll
            for (int c = 0; c < GF_NCH; ++c) z[k][c] = 0.f;
        }
    }
}
// L2 eviction policies: rows that will be re-read nine iterations later are loaded "evict last", their final
// read "evict first" (with ~1200 resident warps the 9-row history of three maps is ~65 MB; plain LRU loses it to
// the streaming traffic and the re-read goes to DRAM: ncu showed 3.7 GB read for 1.9 GB of input)
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void ld_grp_hint(const float* __restrict__ row, int x, int W, float (&z)[4][GF_NCH], uint64_t pol) {
    static_assert(GF_NCH == 4, "float4 per column");
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        if (xx >= 0 && xx < W) {
            asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                         : "=f"(z[k][0]), "=f"(z[k][1]), "=f"(z[k][2]), "=f"(z[k][3])
                         : "l"(row + (size_t)xx * 4), "l"(pol));
        } else {
            z[k][0] = z[k][1] = z[k][2] = z[k][3] = 0.f;
        }
    }
}
__device__ __forceinline__ void st_grp(float* __restrict__ row, int xx, const float (&v)[GF_NCH]) {
    if (GF_NCH == 2) *reinterpret_cast<float2*>(row + (size_t)xx * 4) = make_float2(v[0], v[GF_NCH - 1]);
    else *reinterpret_cast<float4*>(row + (size_t)xx * 4) = make_float4(v[0], v[1 % GF_NCH], v[2 % GF_NCH], v[3 % GF_NCH]);
}

__device__ __forceinline__ void gf_ab(float mz, float cov, float mx, float i1, float i2,
                                      float& A1, float& b1, float& A2, float& b2) {
    A1 = __fmul_rn(cov, i1); b1 = __fmaf_rn(-A1, mx, mz);
    A2 = __fmul_rn(cov, i2); b2 = __fmaf_rn(-A2, mx, mz);
}

// work item -> (image, channel group, strip, chunk); returns false past the end
struct GfItem { int b, grp, x0, y0, rows; size_t qoff; int coff; bool live; };
__device__ __forceinline__ bool gf_item(int item, int nitems, int P, int Q, int nstrips, int nchunks, int outw, int RC,
                                        int H, int W, GfItem& it) {
    // No early return for a surplus warp: it would make the rest of the kernel "possibly divergent" for the
    // compiler (WARPSYNC + collective brackets around every shuffle).  The surplus warp of an odd item count
    // recomputes the last item with all of its stores masked off (it.live == false).
    it.live = item < nitems;
    if (!it.live) item = nitems - 1;
    it.grp = item % P; item /= P;
    const int strip = item % nstrips; item /= nstrips;
    const int chunk = item % nchunks;
    it.b = item / nchunks;
    it.x0 = strip * outw; it.y0 = chunk * RC;
    it.rows = min(RC, H - it.y0);
    it.coff = (it.grp % GF_SUBS) * GF_NCH;
    it.qoff = (((size_t)it.b * Q + it.grp / GF_SUBS) * (size_t)H * W) * 4 + it.coff;   // float offset of (b, quad, 0, 0, coff)
    return true;
}

// MODE 0: forward, writes LF_1e-3 / LF_1e-4.
// MODE 1 (adjoint, "direct" guide term): lf1/lf2 are the INCOMING gradients w.r.t. the two LF maps; writes
//         gxd[grp][b][y][x] = sum_{c in group} sum_e gLF_e,c * mean_A_e,c  (d LF / d guide at fixed mean_A, mean_b).
template <bool VEC, int MODE>
__global__ void __launch_bounds__(GM_WPC * 32)
gf_forward_march_kernel(const float* __restrict__ feat, const float* __restrict__ guide,
                        const float* __restrict__ stats, float* __restrict__ lf1, float* __restrict__ lf2,
                        float* __restrict__ gxd,
                        int P, int Q, int B, int H, int W, int RC, int nstrips, int nchunks, int nitems) {
    extern __shared__ float4 gm_ring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GfItem it;
    gf_item(blockIdx.x * GM_WPC + warp, nitems, P, Q, nstrips, nchunks, GM_OUTW, RC, H, W, it);
    const int b = it.b, x0 = it.x0, y0 = it.y0, rows = it.rows;
    float4* ring = gm_ring + warp * GM_RING_F4 + lane;          // element (slot, i) at [(slot * 2 * GF_NCH + i) * 32]

    const int xr = x0 - 8 + 4 * lane, xs = xr + 4, xo = xr + 8;
    const size_t plane = (size_t)H * W;
    const float* zp = feat + it.qoff;
    const float* gp = guide + (size_t)b * plane;
    const float* mxp = stats + (size_t)b * plane;
    const float* i1p = stats + ((size_t)B + b) * plane;
    const float* i2p = stats + ((size_t)2 * B + b) * plane;
    float* o1 = lf1 + it.qoff;
    float* o2 = lf2 + it.qoff;

    float cs[4], co[4];            // 1 / clipped window width of the level-1 / output columns (0 = column unused)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cs[k] = (xs + k >= 0 && xs + k < W) ? __frcp_rn(win_count(xs + k, W)) : 0.f;
        co[k] = (it.live && xo + k < W && 4 * lane + k < GM_OUTW) ? __frcp_rn(win_count(xo + k, W)) : 0.f;
    }

    float Sz[4][GF_NCH], Sgz[4][GF_NCH], SA1[4][GF_NCH], Sb1[4][GF_NCH], SA2[4][GF_NCH], Sb2[4][GF_NCH];   // [column][channel]
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < GF_NCH; ++c) { Sz[k][c] = Sgz[k][c] = SA1[k][c] = Sb1[k][c] = SA2[k][c] = Sb2[k][c] = 0.f; }

    // raw rows are fetched one iteration ahead: zn/gn = entering row, zq/gq = leaving row of iteration t
    float zn[4][GF_NCH], zq[4][GF_NCH], gn[4], gq[4];
    {
        const int yr = y0 - 8;
        if (yr >= 0) { ld_grp(zp + (size_t)yr * W * 4, xr, W, zn); ld_cols4<VEC>(gp + (size_t)yr * W, xr, W, gn); }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { gn[k] = 0.f;
#pragma unroll
                for (int c = 0; c < GF_NCH; ++c) zn[k][c] = 0.f; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { gq[k] = 0.f;
#pragma unroll
            for (int c = 0; c < GF_NCH; ++c) zq[k][c] = 0.f; }
    }

    int slot = 0;
    const int nt = rows + 16;
    for (int t = 0; t < nt; ++t) {
        const int yr = y0 - 8 + t;                 // raw row entering the level-1 window
        const int ys = yr - 4;                     // level-1 row completed by it (t >= 8)
        const int yo = yr - 8;                     // output row completed (t >= 16)

        // ---- level-1 vertical running sums: + entering row, - leaving row
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < GF_NCH; ++c) {
                Sz[k][c] += zn[k][c] - zq[k][c];
                Sgz[k][c] = __fmaf_rn(-gq[k], zq[k][c], __fmaf_rn(gn[k], zn[k][c], Sgz[k][c]));   // exact products both times
            }
        // ---- prefetch the rows of iteration t+1 (entering: yr+1, leaving: yr-8)
        {
            const int yn = yr + 1, yl = yr - 8;
            if (t + 1 < nt && yn >= 0 && yn < H) { ld_grp(zp + (size_t)yn * W * 4, xr, W, zn); ld_cols4<VEC>(gp + (size_t)yn * W, xr, W, gn); }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { gn[k] = 0.f;
#pragma unroll
                    for (int c = 0; c < GF_NCH; ++c) zn[k][c] = 0.f; }
            }
            if (t + 1 >= 9 && yl >= 0 && yl < H) { ld_grp(zp + (size_t)yl * W * 4, xr, W, zq); ld_cols4<VEC>(gp + (size_t)yl * W, xr, W, gq); }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { gq[k] = 0.f;
#pragma unroll
                    for (int c = 0; c < GF_NCH; ++c) zq[k][c] = 0.f; }
            }
        }
        if (t < 8) continue;

        // ---- level-1 row ys: statistics, A/b of the entering row; A/b of the leaving row (ys - 9) from the ring
        float mx[4], i1[4], i2[4], mxo[4], i1o[4], i2o[4], rn[4];
        const bool row_in = ys >= 0 && ys < H;
        if (row_in) {
            ld_cols4<VEC>(mxp + (size_t)ys * W, xs, W, mx);
            ld_cols4<VEC>(i1p + (size_t)ys * W, xs, W, i1);
            ld_cols4<VEC>(i2p + (size_t)ys * W, xs, W, i2);
            const float rcy = __frcp_rn(win_count(ys, H));
#pragma unroll
            for (int k = 0; k < 4; ++k) rn[k] = rcy * cs[k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { mx[k] = i1[k] = i2[k] = rn[k] = 0.f; }
        }
        const bool has_old = t >= 17 && ys - 9 >= 0;          // (ys - 9 < H always)
        if (has_old) {
            ld_cols4<VEC>(mxp + (size_t)(ys - 9) * W, xs, W, mxo);
            ld_cols4<VEC>(i1p + (size_t)(ys - 9) * W, xs, W, i1o);
            ld_cols4<VEC>(i2p + (size_t)(ys - 9) * W, xs, W, i2o);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { mxo[k] = i1o[k] = i2o[k] = 0.f; }
        }
        float4* rs = ring + slot * 2 * GF_NCH * 32;
#pragma unroll
        for (int c = 0; c < GF_NCH; ++c) {
            float bz[4], bg[4];
            hsum9c(Sz, c, bz);
            hsum9c(Sgz, c, bg);
            float mz[4], cov[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mz[k] = bz[k] * rn[k];
                cov[k] = __fmaf_rn(-mx[k], mz[k], bg[k] * rn[k]);
            }
            float4 old0 = make_float4(0.f, 0.f, 0.f, 0.f), old1 = old0;
            if (has_old) { old0 = rs[(2 * c) * 32]; old1 = rs[(2 * c + 1) * 32]; }
            rs[(2 * c) * 32] = make_float4(mz[0], cov[0], mz[1], cov[1]);
            rs[(2 * c + 1) * 32] = make_float4(mz[2], cov[2], mz[3], cov[3]);
            const float mzo[4] = {old0.x, old0.z, old1.x, old1.z}, covo[4] = {old0.y, old0.w, old1.y, old1.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float A1, b1, A2, b2, A1o, b1o, A2o, b2o;
                gf_ab(mz[k], cov[k], mx[k], i1[k], i2[k], A1, b1, A2, b2);
                gf_ab(mzo[k], covo[k], mxo[k], i1o[k], i2o[k], A1o, b1o, A2o, b2o);
                SA1[k][c] += A1 - A1o; SA2[k][c] += A2 - A2o;
                if (MODE == 0) { Sb1[k][c] += b1 - b1o; Sb2[k][c] += b2 - b2o; }
            }
        }
        slot = slot == 8 ? 0 : slot + 1;
        if (t < 16) continue;

        // ---- output row yo (always inside the image and the chunk)
        float rno[4];
        {
            const float rcy = __frcp_rn(win_count(yo, H));
#pragma unroll
            for (int k = 0; k < 4; ++k) rno[k] = rcy * co[k];
        }
        if (MODE == 0) {
            float go[4];
            ld_cols4<VEC>(gp + (size_t)yo * W, xo, W, go);
            float r1[4][GF_NCH], r2[4][GF_NCH];
#pragma unroll
            for (int c = 0; c < GF_NCH; ++c) {
                float hA[4], hb[4];
                hsum9c(SA1, c, hA);
                hsum9c(Sb1, c, hb);
#pragma unroll
                for (int k = 0; k < 4; ++k) r1[k][c] = __fmaf_rn(hA[k] * rno[k], go[k], hb[k] * rno[k]);
                hsum9c(SA2, c, hA);
                hsum9c(Sb2, c, hb);
#pragma unroll
                for (int k = 0; k < 4; ++k) r2[k][c] = __fmaf_rn(hA[k] * rno[k], go[k], hb[k] * rno[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (co[k] > 0.f) {
                    st_grp(o1 + (size_t)yo * W * 4, xo + k, r1[k]);
                    st_grp(o2 + (size_t)yo * W * 4, xo + k, r2[k]);
                }
            }
        } else {
            float l1[4][GF_NCH], l2[4][GF_NCH];
            ld_grp(o1 + (size_t)yo * W * 4, xo, W, l1);
            ld_grp(o2 + (size_t)yo * W * 4, xo, W, l2);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < GF_NCH; ++c) {
                float h1[4], h2[4];
                hsum9c(SA1, c, h1);
                hsum9c(SA2, c, h2);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = fmaf(l1[k][c], h1[k] * rno[k], fmaf(l2[k][c], h2[k] * rno[k], acc[k]));
            }
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) ok[k] = co[k] > 0.f;
            st_cols4<VEC>(gxd + ((size_t)it.grp * B + b) * plane + (size_t)yo * W, xo, acc, ok);
        }
    }
}

// ------------------------------------------------------------------------------------------
// adjoint, marching formulation (same warp-per-item structure as the forward; three passes):
//   pass A (gf_adjoint_level2_kernel): forward level-1 statistics + adjoint of the level-2 boxes
//       gA'_e = box(gLF_e g / N), gb_e = box(gLF_e / N), then the pointwise chain rule through
//       A_e = cov/(var+eps_e), b_e = mean_z - A_e mean_g; writes gcov/N and gmean_z/N per channel
//       (two maps) and this channel group's share of gvar/N and gmean_g/N (two planes per group).
//   pass B (gf_forward_march_kernel<., 1>): the direct guide term sum gLF_e mean_A_e.
//   pass C (gf_adjoint_level1_kernel): adjoint of the level-1 boxes: g_z = box(gcov/N) g + box(gmean_z/N),
//       g_guide(group) = sum_c box(gcov/N) z_c + 2 g box(gvar/N) + box(gmean_g/N) + direct term.
// Formulas: SURVEY.md 8a (checked there against autograd in fp64).
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(64)
gf_adjoint_level2_kernel(const float* __restrict__ feat, const float* __restrict__ guide, const float* __restrict__ stats,
                         const float* __restrict__ glf1, const float* __restrict__ glf2,
                         float* __restrict__ gc, float* __restrict__ gm, float* __restrict__ gvn, float* __restrict__ gmn,
                         int P, int Q, int B, int H, int W, int RC, int nstrips, int nchunks, int nitems) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GfItem it;
    gf_item(blockIdx.x * 2 + warp, nitems, P, Q, nstrips, nchunks, GA_OUTW, RC, H, W, it);
    const int b = it.b, x0 = it.x0, y0 = it.y0, rows = it.rows;
    const int xr = x0 - 4 + 4 * lane, xs = xr + 4;
    const size_t plane = (size_t)H * W;
    const float* zp = feat + it.qoff;
    const float* l1p = glf1 + it.qoff;
    const float* l2p = glf2 + it.qoff;
    const float* gp = guide + (size_t)b * plane;
    const float* mxp = stats + (size_t)b * plane;
    const float* i1p = stats + ((size_t)B + b) * plane;
    const float* i2p = stats + ((size_t)2 * B + b) * plane;
    float* gcp = gc + it.qoff;
    float* gmp = gm + it.qoff;
    float* gvp = gvn + ((size_t)it.grp * B + b) * plane;
    float* gnp = gmn + ((size_t)it.grp * B + b) * plane;

    float cr[4], cs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cr[k] = (xr + k >= 0 && xr + k < W) ? __frcp_rn(win_count(xr + k, W)) : 0.f;      // reciprocal widths, 0 = unused
        cs[k] = (it.live && xs + k < W && 4 * lane + k < GA_OUTW) ? __frcp_rn(win_count(xs + k, W)) : 0.f;
    }
    const uint64_t pol_keep = l2_policy_keep(), pol_drop = l2_policy_drop();
    float Sz[4][GF_NCH], Sgz[4][GF_NCH], P1[4][GF_NCH], P2[4][GF_NCH], P3[4][GF_NCH], P4[4][GF_NCH];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < GF_NCH; ++c) { Sz[k][c] = Sgz[k][c] = P1[k][c] = P2[k][c] = P3[k][c] = P4[k][c] = 0.f; }

    const int nt = rows + 8;
    for (int t = 0; t < nt; ++t) {
        const int yr = y0 - 4 + t;                 // row entering the vertical windows
#pragma unroll
        for (int side = 0; side < 2; ++side) {     // 0: entering row (+), 1: leaving row (-)
            const int y = side == 0 ? yr : yr - 9;
            if (y < 0 || y >= H || (side == 1 && t < 9)) continue;
            const float sg = side == 0 ? 1.f : -1.f;
            float z[4][GF_NCH], l1[4][GF_NCH], l2[4][GF_NCH], g[4];
            const uint64_t pol = side == 0 ? pol_keep : pol_drop;
            ld_grp_hint(zp + (size_t)y * W * 4, xr, W, z, pol);
            ld_grp_hint(l1p + (size_t)y * W * 4, xr, W, l1, pol);
            ld_grp_hint(l2p + (size_t)y * W * 4, xr, W, l2, pol);
            ld_cols4<VEC>(gp + (size_t)y * W, xr, W, g);
            const float rcy = __frcp_rn(win_count(y, H));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float rn = rcy * cr[k];
                const float w1 = sg * rn, wg = sg * __fmul_rn(g[k], rn), sgk = sg * g[k];
#pragma unroll
                for (int c = 0; c < GF_NCH; ++c) {
                    Sz[k][c] = fmaf(sg, z[k][c], Sz[k][c]);
                    Sgz[k][c] = fmaf(sgk, z[k][c], Sgz[k][c]);
                    P1[k][c] = fmaf(wg, l1[k][c], P1[k][c]);
                    P2[k][c] = fmaf(w1, l1[k][c], P2[k][c]);
                    P3[k][c] = fmaf(wg, l2[k][c], P3[k][c]);
                    P4[k][c] = fmaf(w1, l2[k][c], P4[k][c]);
                }
            }
        }
        if (t < 8) continue;
        const int ys = yr - 4;                     // level-1 row, inside the chunk and the image
        float mx[4], i1[4], i2[4], rn[4];
        ld_cols4<VEC>(mxp + (size_t)ys * W, xs, W, mx);
        ld_cols4<VEC>(i1p + (size_t)ys * W, xs, W, i1);
        ld_cols4<VEC>(i2p + (size_t)ys * W, xs, W, i2);
        {
            const float rcy = __frcp_rn(win_count(ys, H));
#pragma unroll
            for (int k = 0; k < 4; ++k) rn[k] = rcy * cs[k];
        }
        float oc[4][GF_NCH], om[4][GF_NCH], GV[4] = {0.f, 0.f, 0.f, 0.f}, GM[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < GF_NCH; ++c) {
            float bz[4], bg[4], pa1[4], pb1[4], pa2[4], pb2[4];
            hsum9c(Sz, c, bz);
            hsum9c(Sgz, c, bg);
            hsum9c(P1, c, pa1);
            hsum9c(P2, c, pb1);
            hsum9c(P3, c, pa2);
            hsum9c(P4, c, pb2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float mz = bz[k] * rn[k];
                const float cov = fmaf(-mx[k], mz, bg[k] * rn[k]);
                const float A1 = cov * i1[k], A2 = cov * i2[k];
                const float gA1 = fmaf(-pb1[k], mx[k], pa1[k]), gA2 = fmaf(-pb2[k], mx[k], pa2[k]);
                const float gcov = gA1 * i1[k] + gA2 * i2[k];
                const float gmz = pb1[k] + pb2[k] - gcov * mx[k];
                GV[k] -= gA1 * A1 * i1[k] + gA2 * A2 * i2[k];
                GM[k] -= pb1[k] * A1 + pb2[k] * A2 + gcov * mz;
                oc[k][c] = gcov * rn[k];
                om[k][c] = gmz * rn[k];
            }
        }
        float ov[4], on[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ok[k] = cs[k] > 0.f;
            ov[k] = GV[k] * rn[k];
            on[k] = (GM[k] - 2.f * mx[k] * GV[k]) * rn[k];
            if (ok[k]) {
                st_grp(gcp + (size_t)ys * W * 4, xs + k, oc[k]);
                st_grp(gmp + (size_t)ys * W * 4, xs + k, om[k]);
            }
        }
        st_cols4<VEC>(gvp + (size_t)ys * W, xs, ov, ok);
        st_cols4<VEC>(gnp + (size_t)ys * W, xs, on, ok);
    }
}

template <bool VEC>
__global__ void __launch_bounds__(64)
gf_adjoint_level1_kernel(const float* __restrict__ feat, const float* __restrict__ guide,
                         const float* __restrict__ gc, const float* __restrict__ gm,
                         const float* __restrict__ gvn, const float* __restrict__ gmn,
                         float* __restrict__ gfeat, float* __restrict__ gres_partial,
                         int P, int Q, int B, int H, int W, int RC, int nstrips, int nchunks, int nitems) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GfItem it;
    gf_item(blockIdx.x * 2 + warp, nitems, P, Q, nstrips, nchunks, GA_OUTW, RC, H, W, it);
    const int b = it.b, x0 = it.x0, y0 = it.y0, rows = it.rows;
    const int xs = x0 - 4 + 4 * lane, xo = xs + 4;
    const size_t plane = (size_t)H * W;
    const float* zp = feat + it.qoff;
    const float* gcp = gc + it.qoff;
    const float* gmp = gm + it.qoff;
    const float* gp = guide + (size_t)b * plane;
    const float* gvp = gvn + ((size_t)it.grp * B + b) * plane;
    const float* gnp = gmn + ((size_t)it.grp * B + b) * plane;
    float* gzp = gfeat + it.qoff;
    float* gxp = gres_partial + ((size_t)it.grp * B + b) * plane;

    bool ok_o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ok_o[k] = it.live && xo + k < W && 4 * lane + k < GA_OUTW;
    float Sc[4][GF_NCH], Sm[4][GF_NCH], SV[4], SM[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        SV[k] = SM[k] = 0.f;
#pragma unroll
        for (int c = 0; c < GF_NCH; ++c) Sc[k][c] = Sm[k][c] = 0.f;
    }
    // (A shared-memory ring for the leaving row of gcov/N, gmean_z/N halves this kernel's DRAM reads — 4.0 -> 2.5 GB —
    // but caps it at 6 warps per SM and was measured slower, 1.25 vs 0.98 ms; capping the occupancy at 8 warps per SM so
    // that the history fits L2 was slower still, 1.33 ms.  L2 eviction hints are used instead: 0.96 ms.)
    const uint64_t pol_keep = l2_policy_keep(), pol_drop = l2_policy_drop();
    const int nt = rows + 8;
    for (int t = 0; t < nt; ++t) {
        const int ys = y0 - 4 + t;                 // level-1 row entering the vertical window
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int y = side == 0 ? ys : ys - 9;
            if (y < 0 || y >= H || (side == 1 && t < 9)) continue;
            const float sg = side == 0 ? 1.f : -1.f;
            const uint64_t pol = side == 0 ? pol_keep : pol_drop;
            float c4[4][GF_NCH], m4[4][GF_NCH], v[4], n[4];
            ld_grp_hint(gcp + (size_t)y * W * 4, xs, W, c4, pol);
            ld_grp_hint(gmp + (size_t)y * W * 4, xs, W, m4, pol);
            ld_cols4<VEC>(gvp + (size_t)y * W, xs, W, v);
            ld_cols4<VEC>(gnp + (size_t)y * W, xs, W, n);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int c = 0; c < GF_NCH; ++c) {
                    Sc[k][c] = fmaf(sg, c4[k][c], Sc[k][c]);
                    Sm[k][c] = fmaf(sg, m4[k][c], Sm[k][c]);
                }
                SV[k] = fmaf(sg, v[k], SV[k]);
                SM[k] = fmaf(sg, n[k], SM[k]);
            }
        }
        if (t < 8) continue;
        const int yo = ys - 4;                     // output row, inside the chunk and the image
        float z[4][GF_NCH], g[4], gxd[4], bv[4], bm[4];
        ld_grp(zp + (size_t)yo * W * 4, xo, W, z);
        ld_cols4<VEC>(gp + (size_t)yo * W, xo, W, g);
        ld_cols4<VEC>(gxp + (size_t)yo * W, xo, W, gxd);
        hsum9(SV, bv);
        hsum9(SM, bm);
        float gx[4], gz[4][GF_NCH];
#pragma unroll
        for (int k = 0; k < 4; ++k) gx[k] = fmaf(2.f * g[k], bv[k], bm[k]) + gxd[k];
#pragma unroll
        for (int c = 0; c < GF_NCH; ++c) {
            float tt[4], uu[4];
            hsum9c(Sc, c, tt);
            hsum9c(Sm, c, uu);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                gz[k][c] = fmaf(tt[k], g[k], uu[k]);
                gx[k] = fmaf(tt[k], z[k][c], gx[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (ok_o[k]) st_grp(gzp + (size_t)yo * W * 4, xo + k, gz[k]);
        st_cols4<VEC>(gxp + (size_t)yo * W, xo, gx, ok_o);
    }
}

}  // namespace paif

using namespace paif;

static int gf_march_attr() {
    static unsigned long long done = 0;
    int dev;
    if (!attr_needed(done, &dev)) return 0;
    cudaError_t e = cudaFuncSetAttribute(gf_forward_march_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gf_forward_march_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gf_forward_march_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gf_forward_march_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM);
    if (e != cudaSuccess) { set_error("gf smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_mark(done, dev);
    return 0;
}

// Row chunks: ~120 rows when there is plenty of work (the running sums restart per chunk, which bounds their
// rounding drift; 8-16 halo rows of recompute each), shorter (down to 24 rows) when a small batch would otherwise
// leave the GPU with fewer than ~6 warps per SM of these latency-bound marching kernels.  (One 480-row chunk per
// image was measured 8 % slower than four 120-row chunks at batch 16 despite 10 % less halo work.)
static void gf_chunks(int H, long long items_per_chunk, int* nchunks, int* RC) {
    int n = H <= 160 ? 1 : (H + 60) / 120;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = 6LL * sms;
    if (items_per_chunk * n < want) {
        long long m = (want + items_per_chunk - 1) / items_per_chunk;
        const int max_chunks = H / 24 > 1 ? H / 24 : 1;
        n = (int)(m < max_chunks ? m : max_chunks);
    }
    *RC = cdiv(H, n);
    *nchunks = cdiv(H, *RC);
}

extern "C" int paif_gf_guide_stats(const float* residue, float* stats, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(residue && stats, "null pointer");
    PAIF_REQUIRE(B > 0 && B <= 65535, "B out of range");
    PAIF_REQUIRE(H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    dim3 grid(cdiv(W, 32), cdiv(H, 32), B);
    gf_guide_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(residue, stats, B, H, W);
    return check_launch("paif_gf_guide_stats");
}

extern "C" int paif_gf_decomp_forward(const float* feat, const float* residue, const float* stats,
                                      float* lf1, float* lf2, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && residue && stats && lf1 && lf2, "null pointer");
    PAIF_REQUIRE(C > 0 && C % 4 == 0, "C must be a multiple of 4");
    PAIF_REQUIRE(H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    const int Q = C / 4, P = C / GF_NCH;
    const int nstrips = cdiv(W, GM_OUTW);
    int nchunks, RC;
    gf_chunks(H, (long long)B * P * nstrips, &nchunks, &RC);
    const long long nitems = (long long)B * P * nstrips * nchunks;
    PAIF_REQUIRE(nitems < (1ll << 30), "problem too large");
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(residue) | reinterpret_cast<uintptr_t>(stats)) % 16 == 0);
    if (int r = gf_march_attr()) return r;
    const int grid = (int)((nitems + GM_WPC - 1) / GM_WPC);
    if (vec)
        gf_forward_march_kernel<true, 0><<<grid, GM_WPC * 32, GM_SMEM, (cudaStream_t)stream>>>(
            feat, residue, stats, lf1, lf2, nullptr, P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
    else
        gf_forward_march_kernel<false, 0><<<grid, GM_WPC * 32, GM_SMEM, (cudaStream_t)stream>>>(
            feat, residue, stats, lf1, lf2, nullptr, P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
    return check_launch("paif_gf_decomp_forward");
}

extern "C" int paif_gf_guide_parts(int C) { return C / GF_NCH; }

extern "C" long long paif_gf_backward_work_floats(int C, int B, int H, int W) {
    // two C-channel maps (gcov/N, gmean_z/N) + two planes per channel group (gvar/N, gmean_g/N)
    return (long long)B * H * W * (2LL * C + 2LL * (C / GF_NCH));
}

// Direct guide term of the adjoint from the fused forward's saved mean2(A') (paif_gf_mix_forward_save):
//   d out_o / d guide (at fixed window statistics) = mean2(A'_o)  =>  gxd[q][b][pixel] = sum_{o in quad q} gx_o * mean2(A'_o).
// gx, ma: fp32 C4 maps [B][Q][H*W][4]; gxd: [Q][B][H*W] (the layout pass C reads).
__global__ void __launch_bounds__(256)
gf_direct_ma_kernel(const float4* __restrict__ gx, const float4* __restrict__ ma, float* __restrict__ gxd, int Q, int B, long long npix) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= npix) return;
    const int q = blockIdx.y, b = blockIdx.z;
    const size_t src = ((size_t)b * Q + q) * (size_t)npix + (size_t)i;
    const float4 g = __ldg(gx + src), m = __ldg(ma + src);
    gxd[((size_t)q * B + b) * (size_t)npix + (size_t)i] = fmaf(g.x, m.x, fmaf(g.y, m.y, fmaf(g.z, m.z, g.w * m.w)));
}

// `gx` / `mean_a` non-null: the direct guide term comes from the fused forward's saved mean2(A') (pass B') instead of a
// forward recompute (pass B)
static int gf_decomp_backward_impl(const float* feat, const float* residue, const float* stats,
                                   const float* glf1, const float* glf2, const float* gx, const float* mean_a,
                                   float* gfeat, float* gres_partial, float* work,
                                   int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && residue && stats && glf1 && glf2 && gfeat && gres_partial && work, "null pointer");
    PAIF_REQUIRE(C > 0 && C % 4 == 0, "C must be a multiple of 4");
    PAIF_REQUIRE(H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    const int Q = C / 4, P = C / GF_NCH;
    const size_t map = (size_t)B * C * H * W, pplanes = (size_t)P * B * H * W;
    float* gc = work;
    float* gm = work + map;
    float* gvn = work + 2 * map;
    float* gmn = gvn + pplanes;
    int nchunks, RC;
    gf_chunks(H, (long long)B * P * cdiv(W, GA_OUTW), &nchunks, &RC);
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(residue) | reinterpret_cast<uintptr_t>(stats) |
                                       reinterpret_cast<uintptr_t>(work) | reinterpret_cast<uintptr_t>(gres_partial)) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = gf_march_attr()) return r;
    {   // pass A: adjoint of the level-2 boxes + pointwise chain rule
        const int nstrips = cdiv(W, GA_OUTW);
        const long long nitems = (long long)B * P * nstrips * nchunks;
        PAIF_REQUIRE(nitems < (1ll << 30), "problem too large");
        const int grid = (int)((nitems + 1) / 2);
        if (vec) gf_adjoint_level2_kernel<true><<<grid, 64, 0, st>>>(feat, residue, stats, glf1, glf2, gc, gm, gvn, gmn,
                                                                    P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
        else gf_adjoint_level2_kernel<false><<<grid, 64, 0, st>>>(feat, residue, stats, glf1, glf2, gc, gm, gvn, gmn,
                                                                  P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
        if (int r = check_launch("paif_gf_decomp_backward(level 2)")) return r;
    }
    if (mean_a) {   // pass B': direct guide term from the saved mean2(A'): gxd[q] = sum_{o in quad q} gx_o * mean2(A'_o)
        PAIF_REQUIRE(gx && C == 32, "saved mean2(A') needs the fused decomposition (C = 32)");
        PAIF_REQUIRE(B <= 65535, "B out of range");
        const long long npix = (long long)H * W;
        dim3 grid((unsigned)((npix + 255) / 256), Q, B);
        gf_direct_ma_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(gx), reinterpret_cast<const float4*>(mean_a),
                                                  gres_partial, Q, B, npix);
        if (int r = check_launch("paif_gf_decomp_backward(direct term, saved)")) return r;
    } else
    {   // pass B: direct guide term (forward recompute of mean_A)
        const int nstrips = cdiv(W, GM_OUTW);
        const long long nitems = (long long)B * P * nstrips * nchunks;
        const int grid = (int)((nitems + GM_WPC - 1) / GM_WPC);
        float* l1 = const_cast<float*>(glf1);
        float* l2 = const_cast<float*>(glf2);
        if (vec) gf_forward_march_kernel<true, 1><<<grid, GM_WPC * 32, GM_SMEM, st>>>(feat, residue, stats, l1, l2, gres_partial,
                                                                                    P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
        else gf_forward_march_kernel<false, 1><<<grid, GM_WPC * 32, GM_SMEM, st>>>(feat, residue, stats, l1, l2, gres_partial,
                                                                                  P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
        if (int r = check_launch("paif_gf_decomp_backward(direct term)")) return r;
    }
    {   // pass C: adjoint of the level-1 boxes, final gradients
        const int nstrips = cdiv(W, GA_OUTW);
        const long long nitems = (long long)B * P * nstrips * nchunks;
        const int grid = (int)((nitems + 1) / 2);
        if (vec) gf_adjoint_level1_kernel<true><<<grid, 64, 0, st>>>(feat, residue, gc, gm, gvn, gmn, gfeat, gres_partial,
                                                                    P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
        else gf_adjoint_level1_kernel<false><<<grid, 64, 0, st>>>(feat, residue, gc, gm, gvn, gmn, gfeat, gres_partial,
                                                                  P, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
    }
    return check_launch("paif_gf_decomp_backward");
}

extern "C" int paif_gf_decomp_backward(const float* feat, const float* residue, const float* stats,
                                       const float* glf1, const float* glf2,
                                       float* gfeat, float* gres_partial, float* work,
                                       int C, int B, int H, int W, void* stream) {
    return gf_decomp_backward_impl(feat, residue, stats, glf1, glf2, nullptr, nullptr, gfeat, gres_partial, work, C, B, H, W, stream);
}

extern "C" int paif_gf_decomp_backward_saved(const float* feat, const float* residue, const float* stats,
                                             const float* glf1, const float* glf2, const float* gx, const float* mean_a,
                                             float* gfeat, float* gres_partial, float* work,
                                             int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(gx && mean_a, "null pointer");
    PAIF_REQUIRE(((reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(mean_a)) & 15) == 0, "pointers must be 16-byte aligned");
    return gf_decomp_backward_impl(feat, residue, stats, glf1, glf2, gx, mean_a, gfeat, gres_partial, work, C, B, H, W, stream);
}
