// Guided-filter decomposition (Cell_Decom.decomposition, core/model_fusion_auto.py:522-535, with
// the un-vendored guided_filter_pytorch GuidedFilter(r=4, eps)) and its adjoint.
//
//   N = box(1); mx = box(g)/N; var = box(g*g)/N - mx^2
//   per channel c, eps in {1e-3,1e-4}:  mz = box(z_c)/N; cov = box(g z_c)/N - mx mz
//       A = cov/(var+eps); b = mz - A mx;  LF_eps,c = box(A)/N * g + box(b)/N
//   box = 9x9 window sum clipped to the image.
//
// One CTA = one 32x32 output tile of one channel quad of one image.  Two box-filter levels =
// halo 4+4: the 48x48 input region is staged in shared memory, every box filter is two separable
// 9-tap passes through shared memory, 4 outputs per thread, written without running-sum
// subtraction so no cancellation error is introduced (the statistics stay fp32 throughout).
#include "common.cuh"

namespace paif {

constexpr int GF_T = 32;            // output tile
constexpr int GF_R8 = GF_T + 16;    // 48: region with halo 8
constexpr int GF_R4 = GF_T + 8;     // 40: region with halo 4
constexpr int GF_NT = 512;
constexpr float GF_EPS1 = 0.001f, GF_EPS2 = 0.0001f;

// horizontal 9-tap sums: dst[r][c] = sum_{k<9} src(r, c+k); 4 outputs per work item.
template <class Ld4>
__device__ __forceinline__ void box_h(Ld4 ld, float* dst, int dp, int rows, int ocols) {
    const int oc4 = ocols / 4;
    for (int it = threadIdx.x; it < rows * oc4; it += GF_NT) {
        const int r = it / oc4, c = (it - r * oc4) * 4;
        const float4 a = ld(r, c), b = ld(r, c + 4), d = ld(r, c + 8);
        const float m = a.w + b.x + b.y + b.z + b.w + d.x;
        float4 o;
        o.x = a.x + a.y + a.z + m;
        o.y = a.y + a.z + m + d.y;
        o.z = a.z + m + d.y + d.z;
        o.w = m + d.y + d.z + d.w;
        *reinterpret_cast<float4*>(dst + r * dp + c) = o;
    }
}

// vertical 9-tap sums of src[rows][sp]: value(r,c) = sum_{k<9} src[r+k][c]; st(r, c, value).
template <class St>
__device__ __forceinline__ void box_v(const float* src, int sp, int orows, int cols, St st) {
    const int or4 = orows / 4;
    for (int it = threadIdx.x; it < or4 * cols; it += GF_NT) {
        const int seg = it / cols, c = it - seg * cols, r = seg * 4;
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = src[(r + k) * sp + c];
        const float m = v[3] + v[4] + v[5] + v[6] + v[7] + v[8];
        st(r + 0, c, v[0] + v[1] + v[2] + m);
        st(r + 1, c, v[1] + v[2] + m + v[9]);
        st(r + 2, c, v[2] + m + v[9] + v[10]);
        st(r + 3, c, m + v[9] + v[10] + v[11]);
    }
}

// clipped-window pixel count along one axis
__device__ __forceinline__ float win_count(int p, int n) {
    const int lo = p - 4 < 0 ? 0 : p - 4, hi = p + 4 > n - 1 ? n - 1 : p + 4;
    return (float)(hi - lo + 1);
}

struct GfSmem {
    float g[GF_R8 * GF_R8];
    float z[4][GF_R8 * GF_R8];
    float tmp[GF_R8 * GF_R4];
    float mx[GF_R4 * GF_R4];
    float iv1[GF_R4 * GF_R4];
    float iv2[GF_R4 * GF_R4];
    float mz[GF_R4 * GF_R4];
    float ab[4][GF_R4 * GF_R4];           // A1, b1, A2, b2 of the current channel
    float out[2][GF_T * GF_T * 4];        // LF1 / LF2, [pixel][4 channels]
};

template <class S>
__device__ __forceinline__ void gf_load_region(S& s, const float* feat, const float* residue,
                                               int b, int q, int Q, int H, int W, int x0, int y0) {
    const size_t plane = (size_t)H * W;
    const float4* zp = reinterpret_cast<const float4*>(feat) + ((size_t)b * Q + q) * plane;
    const float* gp = residue + (size_t)b * plane;
    for (int i = threadIdx.x; i < GF_R8 * GF_R8; i += GF_NT) {
        const int r = i / GF_R8, c = i - r * GF_R8;
        const int y = y0 - 8 + r, x = x0 - 8 + c;
        float gv = 0.f;
        float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) {
            gv = gp[(size_t)y * W + x];
            zv = zp[(size_t)y * W + x];
        }
        s.g[i] = gv;
        s.z[0][i] = zv.x; s.z[1][i] = zv.y; s.z[2][i] = zv.z; s.z[3][i] = zv.w;
    }
}

// mean_g, 1/(var+eps) on the 40x40 region (zero outside the image).
template <class S>
__device__ __forceinline__ void gf_guide_stats(S& s, int H, int W, int x0, int y0) {
    box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&s.g[r * GF_R8 + c]); },
          s.tmp, GF_R4, GF_R8, GF_R4);
    __syncthreads();
    box_v(s.tmp, GF_R4, GF_R4, GF_R4, [&](int r, int c, float v) {
        const int y = y0 - 4 + r, x = x0 - 4 + c;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
        s.mx[r * GF_R4 + c] = in ? __fdiv_rn(v, win_count(y, H) * win_count(x, W)) : 0.f;
    });
    __syncthreads();
    box_h([&](int r, int c) {
              float4 v = *reinterpret_cast<const float4*>(&s.g[r * GF_R8 + c]);
              return make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
          },
          s.tmp, GF_R4, GF_R8, GF_R4);
    __syncthreads();
    box_v(s.tmp, GF_R4, GF_R4, GF_R4, [&](int r, int c, float v) {
        const int y = y0 - 4 + r, x = x0 - 4 + c;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
        float i1 = 0.f, i2 = 0.f;
        if (in) {
            const float m = s.mx[r * GF_R4 + c];
            const float var = __fdiv_rn(v, win_count(y, H) * win_count(x, W)) - m * m;
            i1 = __fdiv_rn(1.f, var + GF_EPS1);
            i2 = __fdiv_rn(1.f, var + GF_EPS2);
        }
        s.iv1[r * GF_R4 + c] = i1;
        s.iv2[r * GF_R4 + c] = i2;
    });
    __syncthreads();
}

// A1,b1,A2,b2 of channel ch on the 40x40 region -> s.ab (zero outside the image); also s.mz.
template <class S>
__device__ __forceinline__ void gf_channel_ab(S& s, int ch, int H, int W, int x0, int y0) {
    const float* z = s.z[ch];
    box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&z[r * GF_R8 + c]); },
          s.tmp, GF_R4, GF_R8, GF_R4);
    __syncthreads();
    box_v(s.tmp, GF_R4, GF_R4, GF_R4, [&](int r, int c, float v) {
        const int y = y0 - 4 + r, x = x0 - 4 + c;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
        s.mz[r * GF_R4 + c] = in ? __fdiv_rn(v, win_count(y, H) * win_count(x, W)) : 0.f;
    });
    __syncthreads();
    box_h([&](int r, int c) {
              const float4 a = *reinterpret_cast<const float4*>(&s.g[r * GF_R8 + c]);
              const float4 v = *reinterpret_cast<const float4*>(&z[r * GF_R8 + c]);
              return make_float4(a.x * v.x, a.y * v.y, a.z * v.z, a.w * v.w);
          },
          s.tmp, GF_R4, GF_R8, GF_R4);
    __syncthreads();
    box_v(s.tmp, GF_R4, GF_R4, GF_R4, [&](int r, int c, float v) {
        const int y = y0 - 4 + r, x = x0 - 4 + c;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
        float A1 = 0.f, b1 = 0.f, A2 = 0.f, b2 = 0.f;
        if (in) {
            const int i = r * GF_R4 + c;
            const float m = s.mx[i], mzv = s.mz[i];
            const float cov = __fdiv_rn(v, win_count(y, H) * win_count(x, W)) - m * mzv;
            A1 = cov * s.iv1[i]; b1 = mzv - A1 * m;
            A2 = cov * s.iv2[i]; b2 = mzv - A2 * m;
        }
        const int i = r * GF_R4 + c;
        s.ab[0][i] = A1; s.ab[1][i] = b1; s.ab[2][i] = A2; s.ab[3][i] = b2;
    });
    __syncthreads();
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
// One WARP = one work item: (image b, channel quad q, 112-column strip, row chunk).  Lane j owns 4
// adjacent columns x 4 channels.  Both box-filter levels are "vertical running sum in registers,
// then horizontal 9-sum by warp shuffles": level 1 runs on the raw rows (z, g*z), level 2 on
// A_e = cov/(var+eps_e), b_e = mean_z - A_e mean_g.  The row leaving a vertical window is
// re-derived (level 1: re-read from global/L2; level 2: recomputed from (mean_z, cov) kept in a
// 9-row shared-memory ring) and subtracted.  The sums restart at every row chunk (16 halo rows),
// which bounds the running-sum rounding drift; there is no __syncthreads anywhere.
// Each horizontal pass shifts the columns a lane owns by +4: raw columns x0-8+4j.., level-1
// columns x0-4+4j.. (valid for 4j+k <= 119), output columns x0+4j.. (valid for 4j+k <= 111).
// ------------------------------------------------------------------------------------------
constexpr int GM_OUTW = 112;               // output columns per warp
constexpr int GM_WPC = 2;                  // warps (work items) per CTA
constexpr int GM_RING_F4 = 9 * 8 * 32;     // float4 per warp: 9 rows x 8 float4 per lane
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

__device__ __forceinline__ void ld_z4(const float4* __restrict__ row, int x, int W, float4 (&z)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
        z[k] = (x + k >= 0 && x + k < W) ? __ldg(row + x + k) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void gf_ab(float mz, float cov, float mx, float i1, float i2,
                                      float& A1, float& b1, float& A2, float& b2) {
    A1 = __fmul_rn(cov, i1); b1 = __fmaf_rn(-A1, mx, mz);
    A2 = __fmul_rn(cov, i2); b2 = __fmaf_rn(-A2, mx, mz);
}

template <bool VEC>
__global__ void __launch_bounds__(GM_WPC * 32)
gf_forward_march_kernel(const float* __restrict__ feat, const float* __restrict__ guide,
                        const float* __restrict__ stats, float* __restrict__ lf1, float* __restrict__ lf2,
                        int Q, int B, int H, int W, int RC, int nstrips, int nchunks, int nitems) {
    extern __shared__ float4 gm_ring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int item = blockIdx.x * GM_WPC + warp;
    if (item >= nitems) return;
    const int q = item % Q; item /= Q;
    const int strip = item % nstrips; item /= nstrips;
    const int chunk = item % nchunks;
    const int b = item / nchunks;
    float4* ring = gm_ring + warp * GM_RING_F4 + lane;          // element (slot, i) at [(slot * 8 + i) * 32]

    const int x0 = strip * GM_OUTW, y0 = chunk * RC;
    const int rows = min(RC, H - y0);
    const int xr = x0 - 8 + 4 * lane, xs = xr + 4, xo = xr + 8;
    const size_t plane = (size_t)H * W;
    const float4* zp = reinterpret_cast<const float4*>(feat) + ((size_t)b * Q + q) * plane;
    const float* gp = guide + (size_t)b * plane;
    const float* mxp = stats + (size_t)b * plane;
    const float* i1p = stats + ((size_t)B + b) * plane;
    const float* i2p = stats + ((size_t)2 * B + b) * plane;
    float4* o1 = reinterpret_cast<float4*>(lf1) + ((size_t)b * Q + q) * plane;
    float4* o2 = reinterpret_cast<float4*>(lf2) + ((size_t)b * Q + q) * plane;

    float cs[4], co[4];            // clipped window widths of the level-1 / output columns (0 = column unused)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cs[k] = (xs + k >= 0 && xs + k < W) ? win_count(xs + k, W) : 0.f;
        co[k] = (xo + k < W && 4 * lane + k < GM_OUTW) ? win_count(xo + k, W) : 0.f;
    }

    float Sz[4][4], Sgz[4][4], SA1[4][4], Sb1[4][4], SA2[4][4], Sb2[4][4];     // [column k][channel c]
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) { Sz[k][c] = Sgz[k][c] = SA1[k][c] = Sb1[k][c] = SA2[k][c] = Sb2[k][c] = 0.f; }

    // raw rows are fetched one iteration ahead: zn/gn = entering row, zq/gq = leaving row of iteration t
    float4 zn[4], zq[4];
    float gn[4], gq[4];
    {
        const int yr = y0 - 8;
        if (yr >= 0) { ld_z4(zp + (size_t)yr * W, xr, W, zn); ld_cols4<VEC>(gp + (size_t)yr * W, xr, W, gn); }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { zn[k] = make_float4(0.f, 0.f, 0.f, 0.f); gn[k] = 0.f; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { zq[k] = make_float4(0.f, 0.f, 0.f, 0.f); gq[k] = 0.f; }
    }

    int slot = 0;
    const int nt = rows + 16;
    for (int t = 0; t < nt; ++t) {
        const int yr = y0 - 8 + t;                 // raw row entering the level-1 window
        const int ys = yr - 4;                     // level-1 row completed by it (t >= 8)
        const int yo = yr - 8;                     // output row completed (t >= 16)

        // ---- level-1 vertical running sums: + entering row, - leaving row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float zc[4] = {zn[k].x, zn[k].y, zn[k].z, zn[k].w};
            const float zo[4] = {zq[k].x, zq[k].y, zq[k].z, zq[k].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                Sz[k][c] += zc[c] - zo[c];
                Sgz[k][c] += __fmul_rn(gn[k], zc[c]) - __fmul_rn(gq[k], zo[c]);
            }
        }
        // ---- prefetch the rows of iteration t+1 (entering: yr+1, leaving: yr-8)
        {
            const int yn = yr + 1, yl = yr - 8;
            if (t + 1 < nt && yn >= 0 && yn < H) { ld_z4(zp + (size_t)yn * W, xr, W, zn); ld_cols4<VEC>(gp + (size_t)yn * W, xr, W, gn); }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { zn[k] = make_float4(0.f, 0.f, 0.f, 0.f); gn[k] = 0.f; }
            }
            if (t + 1 >= 9 && yl >= 0 && yl < H) { ld_z4(zp + (size_t)yl * W, xr, W, zq); ld_cols4<VEC>(gp + (size_t)yl * W, xr, W, gq); }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { zq[k] = make_float4(0.f, 0.f, 0.f, 0.f); gq[k] = 0.f; }
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
            const float cy = win_count(ys, H);
#pragma unroll
            for (int k = 0; k < 4; ++k) rn[k] = cs[k] > 0.f ? __frcp_rn(cy * cs[k]) : 0.f;
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
        float4* rs = ring + slot * 8 * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float a[4], bz[4], bg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = Sz[k][c];
            hsum9(a, bz);
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = Sgz[k][c];
            hsum9(a, bg);
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
                SA1[k][c] += A1 - A1o; Sb1[k][c] += b1 - b1o;
                SA2[k][c] += A2 - A2o; Sb2[k][c] += b2 - b2o;
            }
        }
        slot = slot == 8 ? 0 : slot + 1;
        if (t < 16) continue;

        // ---- output row yo (always inside the image and the chunk)
        float go[4], rno[4];
        ld_cols4<VEC>(gp + (size_t)yo * W, xo, W, go);
        {
            const float cy = win_count(yo, H);
#pragma unroll
            for (int k = 0; k < 4; ++k) rno[k] = co[k] > 0.f ? __frcp_rn(cy * co[k]) : 0.f;
        }
        float r1[4][4], r2[4][4];                  // [column][channel]
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float a[4], hA[4], hb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = SA1[k][c];
            hsum9(a, hA);
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = Sb1[k][c];
            hsum9(a, hb);
#pragma unroll
            for (int k = 0; k < 4; ++k) r1[k][c] = __fmaf_rn(hA[k] * rno[k], go[k], hb[k] * rno[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = SA2[k][c];
            hsum9(a, hA);
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = Sb2[k][c];
            hsum9(a, hb);
#pragma unroll
            for (int k = 0; k < 4; ++k) r2[k][c] = __fmaf_rn(hA[k] * rno[k], go[k], hb[k] * rno[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (co[k] > 0.f) {
                const size_t o = (size_t)yo * W + xo + k;
                o1[o] = make_float4(r1[k][0], r1[k][1], r1[k][2], r1[k][3]);
                o2[o] = make_float4(r2[k][0], r2[k][1], r2[k][2], r2[k][3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// adjoint (SURVEY.md 8a "Guided-filter adjoint"; checked there against autograd in fp64)
// ------------------------------------------------------------------------------------------
struct GfBwdSmem {
    float g[GF_R8 * GF_R8];
    float z[4][GF_R8 * GF_R8];
    float tmp[GF_R8 * GF_R4];
    float mx[GF_R4 * GF_R4];
    float iv1[GF_R4 * GF_R4];
    float iv2[GF_R4 * GF_R4];
    float mz[GF_R4 * GF_R4];
    float ab[4][GF_R4 * GF_R4];      // A1, b1, A2, b2 (forward recompute)
    float gl[2][GF_R8 * GF_R8];      // incoming dL/dLF_eps of the current channel
    float gab[4][GF_R4 * GF_R4];     // dL/dA1', dL/db1, dL/dA2', dL/db2 ; slots 0/1 reused for g_cov/N, g_my/N
    float gvar[GF_R4 * GF_R4];
    float gmx[GF_R4 * GF_R4];
    float gx[GF_T * GF_T];           // guide gradient accumulated over this quad's channels
    float gy[GF_T * GF_T * 4];
};

__device__ __forceinline__ float f4_get(const float4& v, int j) {
    return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

__global__ void __launch_bounds__(GF_NT)
gf_backward_kernel(const float* __restrict__ feat, const float* __restrict__ residue,
                   const float* __restrict__ glf1, const float* __restrict__ glf2,
                   float* __restrict__ gfeat, float* __restrict__ gres_partial, int Q, int B, int H, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GfBwdSmem& s = *reinterpret_cast<GfBwdSmem*>(smem_raw);
    const int x0 = blockIdx.x * GF_T, y0 = blockIdx.y * GF_T;
    const int q = blockIdx.z % Q, b = blockIdx.z / Q;
    const size_t plane = (size_t)H * W;

    gf_load_region(s, feat, residue, b, q, Q, H, W, x0, y0);
    for (int i = threadIdx.x; i < GF_R4 * GF_R4; i += GF_NT) { s.gvar[i] = 0.f; s.gmx[i] = 0.f; }
    for (int i = threadIdx.x; i < GF_T * GF_T; i += GF_NT) s.gx[i] = 0.f;
    __syncthreads();
    gf_guide_stats(s, H, W, x0, y0);

    const float4* g1p = reinterpret_cast<const float4*>(glf1) + ((size_t)b * Q + q) * plane;
    const float4* g2p = reinterpret_cast<const float4*>(glf2) + ((size_t)b * Q + q) * plane;

    for (int ch = 0; ch < 4; ++ch) {
        gf_channel_ab(s, ch, H, W, x0, y0);
        for (int i = threadIdx.x; i < GF_R8 * GF_R8; i += GF_NT) {
            const int r = i / GF_R8, c = i - r * GF_R8;
            const int y = y0 - 8 + r, x = x0 - 8 + c;
            float a = 0.f, d = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                a = f4_get(g1p[(size_t)y * W + x], ch);
                d = f4_get(g2p[(size_t)y * W + x], ch);
            }
            s.gl[0][i] = a; s.gl[1][i] = d;
        }
        __syncthreads();
        // pointwise term sum_e gLF_e * mean_A_e on the tile
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
            const float* src = s.ab[2 * e];
            box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&src[r * GF_R4 + c]); },
                  s.tmp, GF_T, GF_R4, GF_T);
            __syncthreads();
            const float* gl = s.gl[e];
            box_v(s.tmp, GF_T, GF_T, GF_T, [&](int r, int c, float v) {
                const int y = y0 + r, x = x0 + c;
                if (y < H && x < W) {
                    const float mA = __fdiv_rn(v, win_count(y, H) * win_count(x, W));
                    s.gx[r * GF_T + c] += gl[(r + 8) * GF_R8 + c + 8] * mA;
                }
            });
            __syncthreads();
        }
        // dL/dA' = box(gLF x / N), dL/db = box(gLF / N)   (N of the pixel that owns mean_A / mean_b)
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const float* gl = s.gl[k >> 1];
            const bool withx = (k & 1) == 0;
            box_h([&](int r, int c) {
                      const float4 v = *reinterpret_cast<const float4*>(&gl[r * GF_R8 + c]);
                      const float4 xg = *reinterpret_cast<const float4*>(&s.g[r * GF_R8 + c]);
                      const int y = y0 - 8 + r, x = x0 - 8 + c;
                      const float ny = fmaxf(win_count(y, H), 1.f);
                      float4 o;
                      o.x = __fdiv_rn(withx ? v.x * xg.x : v.x, ny * fmaxf(win_count(x + 0, W), 1.f));
                      o.y = __fdiv_rn(withx ? v.y * xg.y : v.y, ny * fmaxf(win_count(x + 1, W), 1.f));
                      o.z = __fdiv_rn(withx ? v.z * xg.z : v.z, ny * fmaxf(win_count(x + 2, W), 1.f));
                      o.w = __fdiv_rn(withx ? v.w * xg.w : v.w, ny * fmaxf(win_count(x + 3, W), 1.f));
                      return o;
                  },
                  s.tmp, GF_R4, GF_R8, GF_R4);
            __syncthreads();
            float* dst = s.gab[k];
            box_v(s.tmp, GF_R4, GF_R4, GF_R4, [&](int r, int c, float v) { dst[r * GF_R4 + c] = v; });
            __syncthreads();
        }
        // pointwise chain rule on the halo-4 region
        for (int i = threadIdx.x; i < GF_R4 * GF_R4; i += GF_NT) {
            const int r = i / GF_R4, c = i - r * GF_R4;
            const int y = y0 - 4 + r, x = x0 - 4 + c;
            float t = 0.f, u = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float mxv = s.mx[i], my = s.mz[i], A1 = s.ab[0][i], A2 = s.ab[2][i];
                const float i1 = s.iv1[i], i2 = s.iv2[i];
                const float gb1 = s.gab[1][i], gb2 = s.gab[3][i];
                const float gA1 = s.gab[0][i] - gb1 * mxv, gA2 = s.gab[2][i] - gb2 * mxv;
                const float gcov = gA1 * i1 + gA2 * i2;
                s.gvar[i] -= gA1 * A1 * i1 + gA2 * A2 * i2;
                const float gmy = gb1 + gb2 - gcov * mxv;
                s.gmx[i] -= gb1 * A1 + gb2 * A2 + gcov * my;
                const float n = win_count(y, H) * win_count(x, W);
                t = __fdiv_rn(gcov, n);
                u = __fdiv_rn(gmy, n);
            }
            s.gab[0][i] = t; s.gab[1][i] = u;
        }
        __syncthreads();
        {
            const float* src = s.gab[0];
            box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&src[r * GF_R4 + c]); },
                  s.tmp, GF_T, GF_R4, GF_T);
            __syncthreads();
            const float* zc = s.z[ch];
            box_v(s.tmp, GF_T, GF_T, GF_T, [&](int r, int c, float v) {
                const int ci = (r + 8) * GF_R8 + c + 8;
                s.gy[(r * GF_T + c) * 4 + ch] = v * s.g[ci];
                s.gx[r * GF_T + c] += v * zc[ci];
            });
            __syncthreads();
            const float* src2 = s.gab[1];
            box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&src2[r * GF_R4 + c]); },
                  s.tmp, GF_T, GF_R4, GF_T);
            __syncthreads();
            box_v(s.tmp, GF_T, GF_T, GF_T, [&](int r, int c, float v) { s.gy[(r * GF_T + c) * 4 + ch] += v; });
            __syncthreads();
        }
    }
    // var = box(x^2)/N - mx^2 ; mx = box(x)/N
    for (int i = threadIdx.x; i < GF_R4 * GF_R4; i += GF_NT) {
        const int r = i / GF_R4, c = i - r * GF_R4;
        const int y = y0 - 4 + r, x = x0 - 4 + c;
        float gv = 0.f, gm = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const float n = win_count(y, H) * win_count(x, W);
            gm = __fdiv_rn(s.gmx[i] - 2.f * s.mx[i] * s.gvar[i], n);
            gv = __fdiv_rn(s.gvar[i], n);
        }
        s.gvar[i] = gv; s.gmx[i] = gm;
    }
    __syncthreads();
    box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&s.gvar[r * GF_R4 + c]); }, s.tmp, GF_T, GF_R4, GF_T);
    __syncthreads();
    box_v(s.tmp, GF_T, GF_T, GF_T, [&](int r, int c, float v) {
        s.gx[r * GF_T + c] += 2.f * s.g[(r + 8) * GF_R8 + c + 8] * v;
    });
    __syncthreads();
    box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&s.gmx[r * GF_R4 + c]); }, s.tmp, GF_T, GF_R4, GF_T);
    __syncthreads();
    box_v(s.tmp, GF_T, GF_T, GF_T, [&](int r, int c, float v) { s.gx[r * GF_T + c] += v; });
    __syncthreads();

    float4* gyp = reinterpret_cast<float4*>(gfeat) + ((size_t)b * Q + q) * plane;
    float* gxp = gres_partial + ((size_t)q * B + b) * plane;
    for (int i = threadIdx.x; i < GF_T * GF_T; i += GF_NT) {
        const int r = i / GF_T, c = i - r * GF_T;
        const int y = y0 + r, x = x0 + c;
        if (y < H && x < W) {
            gyp[(size_t)y * W + x] = *reinterpret_cast<const float4*>(&s.gy[i * 4]);
            gxp[(size_t)y * W + x] = s.gx[i];
        }
    }
}

}  // namespace paif

using namespace paif;

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
    const int Q = C / 4;
    const int nstrips = cdiv(W, GM_OUTW);
    // row chunks of ~120 rows: the running sums restart per chunk (16 halo rows of recompute each)
    const int nchunks = H <= 160 ? 1 : (H + 60) / 120;
    const int RC = cdiv(H, nchunks);
    const long long nitems = (long long)B * Q * nstrips * nchunks;
    PAIF_REQUIRE(nitems < (1ll << 30), "problem too large");
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(residue) | reinterpret_cast<uintptr_t>(stats)) % 16 == 0);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t err = cudaFuncSetAttribute(gf_forward_march_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(gf_forward_march_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM);
        if (err != cudaSuccess) { set_error("gf smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        attr_done = true;
    }
    const int grid = (int)((nitems + GM_WPC - 1) / GM_WPC);
    if (vec)
        gf_forward_march_kernel<true><<<grid, GM_WPC * 32, GM_SMEM, (cudaStream_t)stream>>>(
            feat, residue, stats, lf1, lf2, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
    else
        gf_forward_march_kernel<false><<<grid, GM_WPC * 32, GM_SMEM, (cudaStream_t)stream>>>(
            feat, residue, stats, lf1, lf2, Q, B, H, W, RC, nstrips, nchunks, (int)nitems);
    return check_launch("paif_gf_decomp_forward");
}

extern "C" int paif_gf_decomp_backward(const float* feat, const float* residue, const float* glf1, const float* glf2,
                                       float* gfeat, float* gres_partial, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && residue && glf1 && glf2 && gfeat && gres_partial, "null pointer");
    PAIF_REQUIRE(C > 0 && C % 4 == 0, "C must be a multiple of 4");
    PAIF_REQUIRE(H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    const int Q = C / 4;
    PAIF_REQUIRE((long long)B * Q <= 65535, "B*C/4 exceeds grid.z");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t err = cudaFuncSetAttribute(gf_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(GfBwdSmem));
        if (err != cudaSuccess) { set_error("gf bwd smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        attr_done = true;
    }
    dim3 grid(cdiv(W, GF_T), cdiv(H, GF_T), B * Q);
    gf_backward_kernel<<<grid, GF_NT, sizeof(GfBwdSmem), (cudaStream_t)stream>>>(feat, residue, glf1, glf2, gfeat,
                                                                               gres_partial, Q, B, H, W);
    return check_launch("paif_gf_decomp_backward");
}
