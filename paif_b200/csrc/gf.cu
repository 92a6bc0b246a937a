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

__global__ void __launch_bounds__(GF_NT)
gf_forward_kernel(const float* __restrict__ feat, const float* __restrict__ residue,
                  float* __restrict__ lf1, float* __restrict__ lf2, int Q, int H, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GfSmem& s = *reinterpret_cast<GfSmem*>(smem_raw);
    const int x0 = blockIdx.x * GF_T, y0 = blockIdx.y * GF_T;
    const int q = blockIdx.z % Q, b = blockIdx.z / Q;

    gf_load_region(s, feat, residue, b, q, Q, H, W, x0, y0);
    __syncthreads();
    gf_guide_stats(s, H, W, x0, y0);

    for (int ch = 0; ch < 4; ++ch) {
        gf_channel_ab(s, ch, H, W, x0, y0);
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {   // A1, b1, A2, b2
            const float* src = s.ab[k];
            box_h([&](int r, int c) { return *reinterpret_cast<const float4*>(&src[r * GF_R4 + c]); },
                  s.tmp, GF_T, GF_R4, GF_T);
            __syncthreads();
            float* o = s.out[k >> 1];
            const bool isA = (k & 1) == 0;
            box_v(s.tmp, GF_T, GF_T, GF_T, [&](int r, int c, float v) {
                const int y = y0 + r, x = x0 + c;
                float m = 0.f;
                if (y < H && x < W) m = __fdiv_rn(v, win_count(y, H) * win_count(x, W));
                const int oi = (r * GF_T + c) * 4 + ch;
                if (isA) o[oi] = m * s.g[(r + 8) * GF_R8 + c + 8];
                else o[oi] += m;
            });
            __syncthreads();
        }
    }
    const size_t plane = (size_t)H * W;
    float4* o1 = reinterpret_cast<float4*>(lf1) + ((size_t)b * Q + q) * plane;
    float4* o2 = reinterpret_cast<float4*>(lf2) + ((size_t)b * Q + q) * plane;
    for (int i = threadIdx.x; i < GF_T * GF_T; i += GF_NT) {
        const int r = i / GF_T, c = i - r * GF_T;
        const int y = y0 + r, x = x0 + c;
        if (y < H && x < W) {
            o1[(size_t)y * W + x] = *reinterpret_cast<const float4*>(&s.out[0][i * 4]);
            o2[(size_t)y * W + x] = *reinterpret_cast<const float4*>(&s.out[1][i * 4]);
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

extern "C" int paif_gf_decomp_forward(const float* feat, const float* residue, float* lf1, float* lf2,
                                      int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && residue && lf1 && lf2, "null pointer");
    PAIF_REQUIRE(C > 0 && C % 4 == 0, "C must be a multiple of 4");
    PAIF_REQUIRE(H > 9 && W > 9, "guided filter needs H, W > 2r+1 = 9");
    const int Q = C / 4;
    PAIF_REQUIRE((long long)B * Q <= 65535, "B*C/4 exceeds grid.z");
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t err = cudaFuncSetAttribute(gf_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(GfSmem));
        if (err != cudaSuccess) { set_error("gf smem attr: %s", cudaGetErrorString(err)); return (int)err; }
        attr_done = true;
    }
    dim3 grid(cdiv(W, GF_T), cdiv(H, GF_T), B * Q);
    gf_forward_kernel<<<grid, GF_NT, sizeof(GfSmem), (cudaStream_t)stream>>>(feat, residue, lf1, lf2, Q, H, W);
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
