// Bandwidth-bound fused kernels of the PAIF fusion path (everything that is not a dense
// convolution or the guided filter).  All work on C4 maps ([B][C/4][H][W][4]); a warp walks x so
// every quad-plane access is a 512-byte coalesced transaction.
#include "common.cuh"

namespace paif {

#define PIX_SETUP()                                            \
    const int x = blockIdx.x * 32 + threadIdx.x;               \
    const int y = blockIdx.y * 8 + threadIdx.y;                \
    const size_t plane = (size_t)H * W;                        \
    const size_t pix = (size_t)y * W + x;

static inline dim3 pix_grid(int W, int H, int Z) { return dim3(cdiv(W, 32), cdiv(H, 8), Z); }

// ------------------------------------------------------------------------------------------
// depthwise dilated conv (DilConv's groups=C BasicConv, operations_m.py:499), one thread per
// (pixel, quad).  w: [C][k*k].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dwconv_kernel(const float* __restrict__ xin, const float* __restrict__ w, int relu_in,
              const float* __restrict__ mask_src, const float* __restrict__ post_res,
              float* __restrict__ out, int Q, int k, int dil, int H, int W) {
    __shared__ float4 sw[49];
    const int q = blockIdx.z % Q, b = blockIdx.z / Q;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int taps = k * k;
    if (tid < taps)
        sw[tid] = make_float4(w[(q * 4 + 0) * taps + tid], w[(q * 4 + 1) * taps + tid],
                              w[(q * 4 + 2) * taps + tid], w[(q * 4 + 3) * taps + tid]);
    __syncthreads();
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float4* xp = reinterpret_cast<const float4*>(xin) + ((size_t)b * Q + q) * plane;
    const int pad = dil * (k - 1) / 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ty = 0; ty < k; ++ty) {
        const int yy = y + ty * dil - pad;
        if (yy < 0 || yy >= H) continue;
        for (int tx = 0; tx < k; ++tx) {
            const int xx = x + tx * dil - pad;
            if (xx < 0 || xx >= W) continue;
            float4 v = xp[(size_t)yy * W + xx];
            if (relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            const float4 ww = sw[ty * k + tx];
            acc.x = fmaf(v.x, ww.x, acc.x); acc.y = fmaf(v.y, ww.y, acc.y);
            acc.z = fmaf(v.z, ww.z, acc.z); acc.w = fmaf(v.w, ww.w, acc.w);
        }
    }
    const size_t off = ((size_t)b * Q + q) * plane + pix;
    if (mask_src) {
        const float4 m = reinterpret_cast<const float4*>(mask_src)[off];
        acc.x = m.x > 0.f ? acc.x : 0.f; acc.y = m.y > 0.f ? acc.y : 0.f;
        acc.z = m.z > 0.f ? acc.z : 0.f; acc.w = m.w > 0.f ? acc.w : 0.f;
    }
    if (post_res) acc = f4_add(acc, reinterpret_cast<const float4*>(post_res)[off]);
    reinterpret_cast<float4*>(out)[off] = acc;
}

// ------------------------------------------------------------------------------------------
// Fused DilConv (operations_m.py:494-506) for C = 32:
//   out = BN(pw1x1(dw_kxk_dil(relu(x)))) + x + r1 + r2          (BN eval-mode scale/shift; r1, r2 optional)
// One thread per pixel keeps the 32 depthwise results' contribution in 32 accumulators; the 1x1
// weights (BN scale folded in, [cin][cout]) and the depthwise taps are broadcast from shared memory.
// HBM-bound design point: x is read once (+L1/L2 halo re-reads), no intermediate map is written.
// ------------------------------------------------------------------------------------------
template <int K, int DIL>
__global__ void __launch_bounds__(256, 2)
dilconv_fused_kernel(const float* __restrict__ xin, const float* __restrict__ dw, const float* __restrict__ pw,
                     const float* __restrict__ ch_scale, const float* __restrict__ ch_shift,
                     const float* __restrict__ r1, const float* __restrict__ r2, float* __restrict__ out,
                     int add_x, int H, int W) {
    constexpr int TAPS = K * K, PAD = DIL * (K - 1) / 2;
    __shared__ __align__(16) float s_pw[32 * 32];      // [cin][cout], scaled by BN
    __shared__ __align__(16) float s_dw[TAPS * 32];    // [tap][channel]
    __shared__ __align__(16) float s_sh[32];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < 1024; i += 256) {
        const int ci = i >> 5, co = i & 31;
        s_pw[i] = pw[co * 32 + ci] * (ch_scale ? ch_scale[co] : 1.f);
    }
    for (int i = tid; i < TAPS * 32; i += 256) {
        const int t = i >> 5, c = i & 31;
        s_dw[i] = dw[c * TAPS + t];
    }
    if (tid < 32) s_sh[tid] = ch_shift ? ch_shift[tid] : 0.f;
    __syncthreads();
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float4* xp = reinterpret_cast<const float4*>(xin) + (size_t)b * 8 * plane;
    // tap offsets (clamped into the image) and validity: the same for every channel quad
    int toff[TAPS];
    float tval[TAPS];
#pragma unroll
    for (int ty = 0; ty < K; ++ty)
#pragma unroll
        for (int tx = 0; tx < K; ++tx) {
            const int yy = y + ty * DIL - PAD, xx = x + tx * DIL - PAD;
            const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
            toff[ty * K + tx] = ok ? yy * W + xx : (int)pix;
            tval[ty * K + tx] = ok ? 1.f : 0.f;
        }
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = s_sh[c];
    float4 v[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) v[t] = __ldg(xp + toff[t]);
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
            const float4 ww = *reinterpret_cast<const float4*>(&s_dw[t * 32 + q * 4]);
            const float m = tval[t];
            t4.x = fmaf(fmaxf(v[t].x, 0.f) * m, ww.x, t4.x); t4.y = fmaf(fmaxf(v[t].y, 0.f) * m, ww.y, t4.y);
            t4.z = fmaf(fmaxf(v[t].z, 0.f) * m, ww.z, t4.z); t4.w = fmaf(fmaxf(v[t].w, 0.f) * m, ww.w, t4.w);
        }
        if (q + 1 < 8) {                                   // next quad's taps are in flight during the 1x1 below
#pragma unroll
            for (int t = 0; t < TAPS; ++t) v[t] = __ldg(xp + (size_t)(q + 1) * plane + toff[t]);
        }
        const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4* wr = reinterpret_cast<const float4*>(&s_pw[(q * 4 + j) * 32]);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w4 = wr[c4];
                acc[c4 * 4 + 0] = fmaf(tv[j], w4.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(tv[j], w4.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(tv[j], w4.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(tv[j], w4.w, acc[c4 * 4 + 3]);
            }
        }
    }
    const size_t base = (size_t)b * 8 * plane + pix;
    float4 rx[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const size_t off = base + q * plane;
        rx[q] = add_x ? __ldg(reinterpret_cast<const float4*>(xin) + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (r1) rx[q] = f4_add(rx[q], __ldg(reinterpret_cast<const float4*>(r1) + off));
        if (r2) rx[q] = f4_add(rx[q], __ldg(reinterpret_cast<const float4*>(r2) + off));
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(out)[base + q * plane] =
            make_float4(acc[q * 4 + 0] + rx[q].x, acc[q * 4 + 1] + rx[q].y, acc[q * 4 + 2] + rx[q].z, acc[q * 4 + 3] + rx[q].w);
}

// ------------------------------------------------------------------------------------------
// 2-arg ChannelPool (core/model_fusion_auto.py:1352-1355)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
channel_pool_kernel(const float* __restrict__ a, const float* __restrict__ v,
                    float* __restrict__ pooled, int Q, int H, int W) {
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float4* ap = reinterpret_cast<const float4*>(a) + (size_t)b * Q * plane + pix;
    const float4* vp = reinterpret_cast<const float4*>(v) + (size_t)b * Q * plane + pix;
    float amax = -INFINITY, asum = 0.f, vmax = -INFINITY, vsum = 0.f;
    for (int q = 0; q < Q; ++q) {
        const float4 t = ap[q * plane];
        amax = fmaxf(fmaxf(amax, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
        asum += (t.x + t.y) + (t.z + t.w);
        const float4 u = vp[q * plane];
        vmax = fmaxf(fmaxf(vmax, fmaxf(u.x, u.y)), fmaxf(u.z, u.w));
        vsum += (u.x + u.y) + (u.z + u.w);
    }
    const float inv = 1.f / (float)(Q * 4);
    reinterpret_cast<float4*>(pooled)[(size_t)b * plane + pix] = make_float4(amax, asum * inv, vmax, vsum * inv);
}

// ------------------------------------------------------------------------------------------
// spatial_attn_layer_M + blend (core/model_fusion_auto.py:1358-1368, :631-632)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spa_blend_kernel(const float* __restrict__ pooled, const float* __restrict__ w, int k,
                 const float* __restrict__ a, const float* __restrict__ v,
                 float* __restrict__ agg, float* __restrict__ scale_out, int Q, int H, int W) {
    __shared__ float4 sw[49];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int taps = k * k;
    if (tid < taps) sw[tid] = make_float4(w[tid], w[taps + tid], w[2 * taps + tid], w[3 * taps + tid]);
    __syncthreads();
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float4* pp = reinterpret_cast<const float4*>(pooled) + (size_t)b * plane;
    const int pad = (k - 1) / 2;
    float acc = 0.f;
    for (int ty = 0; ty < k; ++ty) {
        const int yy = y + ty - pad;
        if (yy < 0 || yy >= H) continue;
        for (int tx = 0; tx < k; ++tx) {
            const int xx = x + tx - pad;
            if (xx < 0 || xx >= W) continue;
            const float4 p = pp[(size_t)yy * W + xx];
            const float4 ww = sw[ty * k + tx];
            acc = fmaf(p.x, ww.x, acc); acc = fmaf(p.y, ww.y, acc);
            acc = fmaf(p.z, ww.z, acc); acc = fmaf(p.w, ww.w, acc);
        }
    }
    const float s = sigmoid_f(acc);
    if (scale_out) scale_out[(size_t)b * plane + pix] = s;
    const float4* ap = reinterpret_cast<const float4*>(a) + (size_t)b * Q * plane + pix;
    const float4* vp = reinterpret_cast<const float4*>(v) + (size_t)b * Q * plane + pix;
    float4* op = reinterpret_cast<float4*>(agg) + (size_t)b * Q * plane + pix;
    const float s1 = 1.f - s;
    for (int q = 0; q < Q; ++q) {
        const float4 t = ap[q * plane], u = vp[q * plane];
        op[q * plane] = make_float4(s * t.x + s1 * u.x, s * t.y + s1 * u.y, s * t.z + s1 * u.z, s * t.w + s1 * u.w);
    }
}

// adjoint, pass 1: gpre = (sum_c G_c (a_c - v_c)) * s (1 - s)
__global__ void __launch_bounds__(256)
spa_blend_bwd_pre_kernel(const float* __restrict__ G, const float* __restrict__ a, const float* __restrict__ v,
                         const float* __restrict__ scale, float* __restrict__ gpre, int Q, int H, int W) {
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const size_t base = (size_t)b * Q * plane + pix;
    float acc = 0.f;
    for (int q = 0; q < Q; ++q) {
        const float4 g = reinterpret_cast<const float4*>(G)[base + q * plane];
        const float4 t = reinterpret_cast<const float4*>(a)[base + q * plane];
        const float4 u = reinterpret_cast<const float4*>(v)[base + q * plane];
        acc = fmaf(g.x, t.x - u.x, acc); acc = fmaf(g.y, t.y - u.y, acc);
        acc = fmaf(g.z, t.z - u.z, acc); acc = fmaf(g.w, t.w - u.w, acc);
    }
    const float s = scale[(size_t)b * plane + pix];
    gpre[(size_t)b * plane + pix] = acc * s * (1.f - s);
}

// adjoint, pass 2: transpose of the k x k conv (1 -> 4), routing through max / mean, blend terms.
__global__ void __launch_bounds__(256)
spa_blend_bwd_kernel(const float* __restrict__ G, const float* __restrict__ a, const float* __restrict__ v,
                     const float* __restrict__ scale, const float* __restrict__ gpre,
                     const float* __restrict__ w, int k,
                     float* __restrict__ ga, float* __restrict__ gv, int Q, int H, int W) {
    __shared__ float4 sw[49];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int taps = k * k;
    if (tid < taps) sw[tid] = make_float4(w[tid], w[taps + tid], w[2 * taps + tid], w[3 * taps + tid]);
    __syncthreads();
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const int pad = (k - 1) / 2;
    const float* gp = gpre + (size_t)b * plane;
    float4 gpool = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ty = 0; ty < k; ++ty) {
        const int yy = y - (ty - pad);
        if (yy < 0 || yy >= H) continue;
        for (int tx = 0; tx < k; ++tx) {
            const int xx = x - (tx - pad);
            if (xx < 0 || xx >= W) continue;
            const float gval = gp[(size_t)yy * W + xx];
            const float4 ww = sw[ty * k + tx];
            gpool.x = fmaf(gval, ww.x, gpool.x); gpool.y = fmaf(gval, ww.y, gpool.y);
            gpool.z = fmaf(gval, ww.z, gpool.z); gpool.w = fmaf(gval, ww.w, gpool.w);
        }
    }
    const size_t base = (size_t)b * Q * plane + pix;
    // arg-max channels (first index on ties, as torch.max(dim))
    int ia = 0, iv = 0;
    float ma = -INFINITY, mv = -INFINITY;
    for (int q = 0; q < Q; ++q) {
        const float4 t = reinterpret_cast<const float4*>(a)[base + q * plane];
        const float4 u = reinterpret_cast<const float4*>(v)[base + q * plane];
        const float tt[4] = {t.x, t.y, t.z, t.w}, uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (tt[j] > ma) { ma = tt[j]; ia = q * 4 + j; }
            if (uu[j] > mv) { mv = uu[j]; iv = q * 4 + j; }
        }
    }
    const float s = scale[(size_t)b * plane + pix], s1 = 1.f - s;
    const float invC = 1.f / (float)(Q * 4);
    const float ameanv = gpool.y * invC, vmeanv = gpool.w * invC;
    for (int q = 0; q < Q; ++q) {
        const float4 g = reinterpret_cast<const float4*>(G)[base + q * plane];
        float ra[4] = {s * g.x + ameanv, s * g.y + ameanv, s * g.z + ameanv, s * g.w + ameanv};
        float rv[4] = {s1 * g.x + vmeanv, s1 * g.y + vmeanv, s1 * g.z + vmeanv, s1 * g.w + vmeanv};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (q * 4 + j == ia) ra[j] += gpool.x;
            if (q * 4 + j == iv) rv[j] += gpool.z;
        }
        reinterpret_cast<float4*>(ga)[base + q * plane] = make_float4(ra[0], ra[1], ra[2], ra[3]);
        reinterpret_cast<float4*>(gv)[base + q * plane] = make_float4(rv[0], rv[1], rv[2], rv[3]);
    }
}

// ------------------------------------------------------------------------------------------
// ECA (operations_m.py:340-393)
// ------------------------------------------------------------------------------------------
__global__ void eca_scale_kernel(const float* __restrict__ partials, int tiles, const float* __restrict__ w1d,
                                 int k, float* __restrict__ e, int C, float inv_hw) {
    extern __shared__ float sm[];   // [C]
    const int b = blockIdx.x, c = threadIdx.x;
    float s = 0.f;
    for (int t = 0; t < tiles; ++t) s += partials[((size_t)b * tiles + t) * C + c];   // fixed order
    sm[c] = s * inv_hw;
    __syncthreads();
    const int pad = (k - 1) / 2;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
        const int cc = c + j - pad;
        if (cc >= 0 && cc < C) acc = fmaf(w1d[j], sm[cc], acc);
    }
    e[(size_t)b * C + c] = sigmoid_f(acc);
}

__global__ void __launch_bounds__(256)
eca_apply_kernel(const float* __restrict__ o, const float* __restrict__ xin, const float* __restrict__ e,
                 const float* __restrict__ slope_p, const float* __restrict__ post_res,
                 float* __restrict__ out, int Q, int H, int W) {
    const int q = blockIdx.z % Q, b = blockIdx.z / Q;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float a = *slope_p;
    const float4 ev = reinterpret_cast<const float4*>(e)[(size_t)b * Q + q];
    const size_t off = ((size_t)b * Q + q) * plane + pix;
    const float4 ov = reinterpret_cast<const float4*>(o)[off];
    const float4 xv = reinterpret_cast<const float4*>(xin)[off];
    float4 r = make_float4(prelu_f(fmaf(ov.x, ev.x, xv.x), a), prelu_f(fmaf(ov.y, ev.y, xv.y), a),
                           prelu_f(fmaf(ov.z, ev.z, xv.z), a), prelu_f(fmaf(ov.w, ev.w, xv.w), a));
    if (post_res) r = f4_add(r, reinterpret_cast<const float4*>(post_res)[off]);
    reinterpret_cast<float4*>(out)[off] = r;
}

// bf16 C8 maps: one thread per (pixel, 8-channel plane)
__global__ void __launch_bounds__(256)
eca_apply_bf16_kernel(const uint4* __restrict__ o, const uint4* __restrict__ xin, const float* __restrict__ e,
                      const float* __restrict__ slope_p, const uint4* __restrict__ post_res,
                      uint4* __restrict__ out, int P, int H, int W) {
    const int p = blockIdx.z % P, b = blockIdx.z / P;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float a = *slope_p;
    const float4 e0 = reinterpret_cast<const float4*>(e)[((size_t)b * P + p) * 2];
    const float4 e1 = reinterpret_cast<const float4*>(e)[((size_t)b * P + p) * 2 + 1];
    const size_t off = ((size_t)b * P + p) * plane + pix;
    float4 o0, o1, x0, x1;
    bf8_unpack(__ldg(o + off), o0, o1);
    bf8_unpack(__ldg(xin + off), x0, x1);
    float4 r0 = make_float4(prelu_f(fmaf(o0.x, e0.x, x0.x), a), prelu_f(fmaf(o0.y, e0.y, x0.y), a),
                            prelu_f(fmaf(o0.z, e0.z, x0.z), a), prelu_f(fmaf(o0.w, e0.w, x0.w), a));
    float4 r1 = make_float4(prelu_f(fmaf(o1.x, e1.x, x1.x), a), prelu_f(fmaf(o1.y, e1.y, x1.y), a),
                            prelu_f(fmaf(o1.z, e1.z, x1.z), a), prelu_f(fmaf(o1.w, e1.w, x1.w), a));
    if (post_res) {
        float4 p0, p1;
        bf8_unpack(__ldg(post_res + off), p0, p1);
        r0 = f4_add(r0, p0);
        r1 = f4_add(r1, p1);
    }
    out[off] = bf8_pack(r0, r1);
}

// backward pass 1: gw = gu * PReLU'(o e + x); per-(b, tile, c) sums of gw * o.  Tile = 32 x 64 px.
constexpr int ECAB_ROWS = 64;
__global__ void __launch_bounds__(256)
eca_bwd_pass1_kernel(const float* __restrict__ gu, const float* __restrict__ o, const float* __restrict__ xin,
                     const float* __restrict__ e, const float* __restrict__ slope_p,
                     float* __restrict__ gw, float* __restrict__ partials, int Q, int H, int W) {
    __shared__ float red[8][32];
    const int b = blockIdx.z;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const size_t plane = (size_t)H * W;
    const float a = *slope_p;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tiles = gridDim.x * gridDim.y, tile = blockIdx.y * gridDim.x + blockIdx.x;
    for (int q = 0; q < Q; ++q) {
        const float4 ev = reinterpret_cast<const float4*>(e)[(size_t)b * Q + q];
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < ECAB_ROWS / 8; ++i) {
            const int y = blockIdx.y * ECAB_ROWS + i * 8 + threadIdx.y;
            if (x < W && y < H) {
                const size_t off = ((size_t)b * Q + q) * plane + (size_t)y * W + x;
                const float4 ov = reinterpret_cast<const float4*>(o)[off];
                const float4 xv = reinterpret_cast<const float4*>(xin)[off];
                const float4 g = reinterpret_cast<const float4*>(gu)[off];
                float4 r;
                r.x = g.x * dprelu_f(fmaf(ov.x, ev.x, xv.x), a);
                r.y = g.y * dprelu_f(fmaf(ov.y, ev.y, xv.y), a);
                r.z = g.z * dprelu_f(fmaf(ov.z, ev.z, xv.z), a);
                r.w = g.w * dprelu_f(fmaf(ov.w, ev.w, xv.w), a);
                reinterpret_cast<float4*>(gw)[off] = r;
                s.x = fmaf(r.x, ov.x, s.x); s.y = fmaf(r.y, ov.y, s.y);
                s.z = fmaf(r.z, ov.z, s.z); s.w = fmaf(r.w, ov.w, s.w);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s.x += __shfl_xor_sync(0xffffffffu, s.x, d); s.y += __shfl_xor_sync(0xffffffffu, s.y, d);
            s.z += __shfl_xor_sync(0xffffffffu, s.z, d); s.w += __shfl_xor_sync(0xffffffffu, s.w, d);
        }
        if (lane == 0) { red[warp][q * 4 + 0] = s.x; red[warp][q * 4 + 1] = s.y; red[warp][q * 4 + 2] = s.z; red[warp][q * 4 + 3] = s.w; }
    }
    __syncthreads();
    const int tid = warp * 32 + lane;
    if (tid < Q * 4) {
        float t = 0.f;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) t += red[wi][tid];
        partials[((size_t)b * tiles + tile) * (Q * 4) + tid] = t;
    }
}

__global__ void eca_bwd_scale_kernel(const float* __restrict__ partials, int tiles, const float* __restrict__ e,
                                     const float* __restrict__ w1d, int k, float* __restrict__ gm,
                                     int C, float inv_hw) {
    extern __shared__ float sm[];   // d[C]
    const int b = blockIdx.x, c = threadIdx.x;
    float s = 0.f;
    for (int t = 0; t < tiles; ++t) s += partials[((size_t)b * tiles + t) * C + c];
    const float ev = e[(size_t)b * C + c];
    sm[c] = s * ev * (1.f - ev);
    __syncthreads();
    const int pad = (k - 1) / 2;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
        const int cc = c - j + pad;
        if (cc >= 0 && cc < C) acc = fmaf(w1d[j], sm[cc], acc);
    }
    gm[(size_t)b * C + c] = acc * inv_hw;
}

__global__ void __launch_bounds__(256)
eca_bwd_pass2_kernel(const float* __restrict__ gw, const float* __restrict__ e, const float* __restrict__ gm,
                     float* __restrict__ go, int Q, int H, int W) {
    const int q = blockIdx.z % Q, b = blockIdx.z / Q;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float4 ev = reinterpret_cast<const float4*>(e)[(size_t)b * Q + q];
    const float4 mv = reinterpret_cast<const float4*>(gm)[(size_t)b * Q + q];
    const size_t off = ((size_t)b * Q + q) * plane + pix;
    const float4 g = reinterpret_cast<const float4*>(gw)[off];
    reinterpret_cast<float4*>(go)[off] =
        make_float4(fmaf(g.x, ev.x, mv.x), fmaf(g.y, ev.y, mv.y), fmaf(g.z, ev.z, mv.z), fmaf(g.w, ev.w, mv.w));
}

// ------------------------------------------------------------------------------------------
// stem_out + tanh as one 5x5 32->1 stencil with 9 border classes (see include/paif_b200.h)
// ------------------------------------------------------------------------------------------
constexpr int OUT_C = 32;
__device__ __forceinline__ int border_class(int p, int n) { return p == 0 ? 0 : (p == n - 1 ? 2 : 1); }

// Forward of the merged stem_out stencil.  Every feature pixel q is read ONCE: its 25 tap products
// d_t(q) = <w_t, f(q)> (interior-class weights) go to shared memory, then each output pixel gathers
// out(p) = sum_t d_t(p + t).  The one-pixel image border (where the second conv's zero padding
// changes the merged weights) is recomputed directly with its class weights.
constexpr int OF_TX = 32, OF_TY = 16;                          // output tile
constexpr int OF_RX = OF_TX + 4, OF_RY = OF_TY + 4;            // feature region (halo 2)
constexpr int OF_NQ = OF_RX * OF_RY;                           // 720
constexpr int OF_SMEM = (25 * OF_NQ + 25 * OUT_C) * 4;

template <bool BF>                                             // BF: feat is a bf16 C8 map
__global__ void __launch_bounds__(256)
out_forward_kernel(const void* __restrict__ feat, const float* __restrict__ wm, const float* __restrict__ slope_p,
                   float* __restrict__ out, float* __restrict__ pre_out, int H, int W) {
    extern __shared__ __align__(16) float of_smem[];
    float* sd = of_smem;                         // [25][OF_NQ]
    float* sw = of_smem + 25 * OF_NQ;            // [25][32] interior class (cls 4)
    const int tid = threadIdx.x;
    for (int i = tid; i < 25 * OUT_C; i += 256) sw[i] = wm[4 * 25 * OUT_C + i];
    __syncthreads();
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * OF_TX, y0 = blockIdx.y * OF_TY;
    const size_t plane = (size_t)H * W;
    const void* fp = img32<BF>(feat, b, plane);
    // two feature pixels per thread and pass (q, q + 256): every weight vector read from shared memory feeds 8 FMAs
    // instead of 4 (the loop is LSU-issue bound on the broadcast weight reads)
    for (int q0 = tid; q0 < OF_NQ; q0 += 512) {
        const int q1 = q0 + 256;
        const bool has1 = q1 < OF_NQ;
        float4 v0[OUT_C / 4], v1[OUT_C / 4];
        {
            const int ry = q0 / OF_RX, rx = q0 - ry * OF_RX;
            const int yy = y0 - 2 + ry, xx = x0 - 2 + rx;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            if (in) ld_px32<BF>(fp, plane, (size_t)yy * W + xx, v0);
            else {
#pragma unroll
                for (int c = 0; c < OUT_C / 4; ++c) v0[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        {
            const int ry = q1 / OF_RX, rx = q1 - ry * OF_RX;
            const int yy = y0 - 2 + ry, xx = x0 - 2 + rx;
            const bool in = has1 && yy >= 0 && yy < H && xx >= 0 && xx < W;
            if (in) ld_px32<BF>(fp, plane, (size_t)yy * W + xx, v1);
            else {
#pragma unroll
                for (int c = 0; c < OUT_C / 4; ++c) v1[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll 5
        for (int t = 0; t < 25; ++t) {
            const float4* wv = reinterpret_cast<const float4*>(sw + t * OUT_C);
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int c = 0; c < OUT_C / 4; c += 2) {
                const float4 w0 = wv[c], w1 = wv[c + 1];
                a0 = fmaf(v0[c].x, w0.x, a0); a0 = fmaf(v0[c].y, w0.y, a0); a0 = fmaf(v0[c].z, w0.z, a0); a0 = fmaf(v0[c].w, w0.w, a0);
                a1 = fmaf(v0[c + 1].x, w1.x, a1); a1 = fmaf(v0[c + 1].y, w1.y, a1); a1 = fmaf(v0[c + 1].z, w1.z, a1); a1 = fmaf(v0[c + 1].w, w1.w, a1);
                b0 = fmaf(v1[c].x, w0.x, b0); b0 = fmaf(v1[c].y, w0.y, b0); b0 = fmaf(v1[c].z, w0.z, b0); b0 = fmaf(v1[c].w, w0.w, b0);
                b1 = fmaf(v1[c + 1].x, w1.x, b1); b1 = fmaf(v1[c + 1].y, w1.y, b1); b1 = fmaf(v1[c + 1].z, w1.z, b1); b1 = fmaf(v1[c + 1].w, w1.w, b1);
            }
            sd[t * OF_NQ + q0] = a0 + a1;
            if (has1) sd[t * OF_NQ + q1] = b0 + b1;
        }
    }
    __syncthreads();
    const float a = __ldg(slope_p);
    for (int o = tid; o < OF_TX * OF_TY; o += 256) {
        const int oy = o / OF_TX, ox = o - oy * OF_TX;
        const int y = y0 + oy, x = x0 + ox;
        if (y >= H || x >= W) continue;
        if (border_class(y, H) != 1 || border_class(x, W) != 1) continue;     // image border: out_border_kernel
        float acc = 0.f;
#pragma unroll
        for (int ty = 0; ty < 5; ++ty)
#pragma unroll
            for (int tx = 0; tx < 5; ++tx) acc += sd[(ty * 5 + tx) * OF_NQ + (oy + ty) * OF_RX + ox + tx];
        const size_t pi = (size_t)b * plane + (size_t)y * W + x;
        if (pre_out) pre_out[pi] = acc;
        out[pi] = tanhf(prelu_f(acc, a));
    }
}

// The one-pixel image border of the merged stem_out stencil: its weights depend on the border class
// (which taps of the second 3x3 conv fall on its zero padding).  2(W + H) - 4 pixels per image.
template <bool BF>
__global__ void __launch_bounds__(128)
out_border_kernel(const void* __restrict__ feat, const float* __restrict__ wm, const float* __restrict__ slope_p,
                  float* __restrict__ out, float* __restrict__ pre_out, int H, int W) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int per = 2 * W + 2 * (H - 2);
    if (i >= per) return;
    int x, y;
    if (i < W) { y = 0; x = i; }
    else if (i < 2 * W) { y = H - 1; x = i - W; }
    else { const int j = i - 2 * W; y = 1 + (j >> 1); x = (j & 1) ? W - 1 : 0; }
    const size_t plane = (size_t)H * W;
    const void* fp = img32<BF>(feat, b, plane);
    const int cls = border_class(y, H) * 3 + border_class(x, W);
    float acc = 0.f;
    for (int ty = 0; ty < 5; ++ty) {
        const int yy = y + ty - 2;
        if (yy < 0 || yy >= H) continue;
        for (int tx = 0; tx < 5; ++tx) {
            const int xx = x + tx - 2;
            if (xx < 0 || xx >= W) continue;
            const float4* wv = reinterpret_cast<const float4*>(wm + ((size_t)cls * 25 + ty * 5 + tx) * OUT_C);
            float4 fv[OUT_C / 4];
            ld_px32<BF>(fp, plane, (size_t)yy * W + xx, fv);
#pragma unroll
            for (int c = 0; c < OUT_C / 4; ++c) {
                const float4 v = fv[c];
                const float4 ww = __ldg(wv + c);
                acc = fmaf(v.x, ww.x, acc); acc = fmaf(v.y, ww.y, acc);
                acc = fmaf(v.z, ww.z, acc); acc = fmaf(v.w, ww.w, acc);
            }
        }
    }
    const size_t pi = (size_t)b * plane + (size_t)y * W + x;
    if (pre_out) pre_out[pi] = acc;
    out[pi] = tanhf(prelu_f(acc, __ldg(slope_p)));
}

__global__ void __launch_bounds__(256)
out_backward_kernel(const float* __restrict__ g, const float* __restrict__ outv, const float* __restrict__ pre,
                    const float* __restrict__ wm, const float* __restrict__ slope_p,
                    float* __restrict__ gfeat, const float* __restrict__ mask_src,
                    const float* __restrict__ mask_slope, float* __restrict__ gfeat_masked, int H, int W) {
    extern __shared__ __align__(16) float swm[];   // [9][25][32]
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < 9 * 25 * OUT_C; i += 256) swm[i] = wm[i];
    __syncthreads();
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const float a = *slope_p;
    float acc[OUT_C];
#pragma unroll
    for (int c = 0; c < OUT_C; ++c) acc[c] = 0.f;
    for (int ty = 0; ty < 5; ++ty) {
        const int yy = y - (ty - 2);          // the output pixel p = q - t that used tap t on q
        if (yy < 0 || yy >= H) continue;
        for (int tx = 0; tx < 5; ++tx) {
            const int xx = x - (tx - 2);
            if (xx < 0 || xx >= W) continue;
            const size_t pi = (size_t)b * plane + (size_t)yy * W + xx;
            const float o = outv[pi];
            const float gs = g[pi] * (1.f - o * o) * dprelu_f(pre[pi], a);
            const int cls = border_class(yy, H) * 3 + border_class(xx, W);
            const float4* wv = reinterpret_cast<const float4*>(swm + ((size_t)cls * 25 + ty * 5 + tx) * OUT_C);
#pragma unroll
            for (int q = 0; q < OUT_C / 4; ++q) {
                const float4 ww = wv[q];
                acc[q * 4 + 0] = fmaf(gs, ww.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(gs, ww.y, acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(gs, ww.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(gs, ww.w, acc[q * 4 + 3]);
            }
        }
    }
    const size_t base = (size_t)b * (OUT_C / 4) * plane + pix;
    const float ma = mask_slope ? *mask_slope : 0.f;
#pragma unroll
    for (int q = 0; q < OUT_C / 4; ++q) {
        const float4 r = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        reinterpret_cast<float4*>(gfeat)[base + q * plane] = r;
        if (gfeat_masked) {
            const float4 m = reinterpret_cast<const float4*>(mask_src)[base + q * plane];
            reinterpret_cast<float4*>(gfeat_masked)[base + q * plane] =
                make_float4(r.x * dprelu_f(m.x, ma), r.y * dprelu_f(m.y, ma), r.z * dprelu_f(m.z, ma), r.w * dprelu_f(m.w, ma));
        }
    }
}

__global__ void __launch_bounds__(256)
mask_scale_kernel(const float* __restrict__ g, const float* __restrict__ mask_src, const float* __restrict__ mask_slope,
                  float scale, float* __restrict__ out, size_t n4) {
    const float ma = mask_slope ? *mask_slope : 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(g)[i];
        if (mask_src) {
            const float4 m = reinterpret_cast<const float4*>(mask_src)[i];
            v.x *= dprelu_f(m.x, ma); v.y *= dprelu_f(m.y, ma); v.z *= dprelu_f(m.z, ma); v.w *= dprelu_f(m.w, ma);
        }
        reinterpret_cast<float4*>(out)[i] = f4_scale(v, scale);
    }
}

__global__ void __launch_bounds__(256)
add_maps_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                float* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = f4_add(reinterpret_cast<const float4*>(a)[i], reinterpret_cast<const float4*>(b)[i]);
        if (c) v = f4_add(v, reinterpret_cast<const float4*>(c)[i]);
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// ------------------------------------------------------------------------------------------
// confusion matrix (robust_test.py:207-211): integer counts, shared-memory histogram per CTA,
// 64-bit integer atomics -> bit-exact regardless of scheduling.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
confusion_kernel(const long long* __restrict__ label, const long long* __restrict__ pred, long long count,
                 int n, unsigned long long* __restrict__ conf) {
    extern __shared__ unsigned int hist[];   // [n*n]
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const long long l = label[i], p = pred[i];
        if (l >= 0 && l < n && p >= 0 && p < n) atomicAdd(&hist[(int)l * n + (int)p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += blockDim.x)
        if (hist[i]) atomicAdd(&conf[i], (unsigned long long)hist[i]);
}

}  // namespace paif

using namespace paif;
#define ST ((cudaStream_t)stream)

// Fused DilConv on bf16 C8 maps: the same thread-per-pixel structure, eight input channels (one 16-byte vector per
// tap) per step; depthwise and 1x1 arithmetic in fp32 registers, one rounding at the store.
template <int K, int DIL>
__global__ void __launch_bounds__(256, 2)
dilconv_fused_bf16_kernel(const uint4* __restrict__ xin, const float* __restrict__ dw, const float* __restrict__ pw,
                          const float* __restrict__ ch_scale, const float* __restrict__ ch_shift,
                          const uint4* __restrict__ r1, const uint4* __restrict__ r2, uint4* __restrict__ out,
                          int add_x, int H, int W) {
    constexpr int TAPS = K * K, PAD = DIL * (K - 1) / 2;
    __shared__ __align__(16) float s_pw[32 * 32];      // [cin][cout], scaled by BN
    __shared__ __align__(16) float s_dw[TAPS * 32];    // [tap][channel]
    __shared__ __align__(16) float s_sh[32];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < 1024; i += 256) {
        const int ci = i >> 5, co = i & 31;
        s_pw[i] = pw[co * 32 + ci] * (ch_scale ? ch_scale[co] : 1.f);
    }
    for (int i = tid; i < TAPS * 32; i += 256) {
        const int t = i >> 5, c = i & 31;
        s_dw[i] = dw[c * TAPS + t];
    }
    if (tid < 32) s_sh[tid] = ch_shift ? ch_shift[tid] : 0.f;
    __syncthreads();
    const int b = blockIdx.z;
    PIX_SETUP();
    if (x >= W || y >= H) return;
    const uint4* xp = xin + (size_t)b * 4 * plane;
    int toff[TAPS];
    float tval[TAPS];
#pragma unroll
    for (int ty = 0; ty < K; ++ty)
#pragma unroll
        for (int tx = 0; tx < K; ++tx) {
            const int yy = y + ty * DIL - PAD, xx = x + tx * DIL - PAD;
            const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
            toff[ty * K + tx] = ok ? yy * W + xx : (int)pix;
            tval[ty * K + tx] = ok ? 1.f : 0.f;
        }
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = s_sh[c];
    uint4 v[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) v[t] = __ldg(xp + toff[t]);
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        float tv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
            float4 lo, hi;
            bf8_unpack(v[t], lo, hi);
            const float4 w0 = *reinterpret_cast<const float4*>(&s_dw[t * 32 + p * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_dw[t * 32 + p * 8 + 4]);
            const float m = tval[t];
            tv[0] = fmaf(fmaxf(lo.x, 0.f) * m, w0.x, tv[0]); tv[1] = fmaf(fmaxf(lo.y, 0.f) * m, w0.y, tv[1]);
            tv[2] = fmaf(fmaxf(lo.z, 0.f) * m, w0.z, tv[2]); tv[3] = fmaf(fmaxf(lo.w, 0.f) * m, w0.w, tv[3]);
            tv[4] = fmaf(fmaxf(hi.x, 0.f) * m, w1.x, tv[4]); tv[5] = fmaf(fmaxf(hi.y, 0.f) * m, w1.y, tv[5]);
            tv[6] = fmaf(fmaxf(hi.z, 0.f) * m, w1.z, tv[6]); tv[7] = fmaf(fmaxf(hi.w, 0.f) * m, w1.w, tv[7]);
        }
        if (p + 1 < 4) {                                   // next plane's taps are in flight during the 1x1 below
#pragma unroll
            for (int t = 0; t < TAPS; ++t) v[t] = __ldg(xp + (size_t)(p + 1) * plane + toff[t]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4* wr = reinterpret_cast<const float4*>(&s_pw[(p * 8 + j) * 32]);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w4 = wr[c4];
                acc[c4 * 4 + 0] = fmaf(tv[j], w4.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(tv[j], w4.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(tv[j], w4.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(tv[j], w4.w, acc[c4 * 4 + 3]);
            }
        }
    }
    const size_t base = (size_t)b * 4 * plane + pix;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const size_t off = base + p * plane;
        float4 lo = make_float4(acc[p * 8 + 0], acc[p * 8 + 1], acc[p * 8 + 2], acc[p * 8 + 3]);
        float4 hi = make_float4(acc[p * 8 + 4], acc[p * 8 + 5], acc[p * 8 + 6], acc[p * 8 + 7]);
        float4 a0, a1;
        if (add_x) { bf8_unpack(__ldg(xin + off), a0, a1); lo = f4_add(lo, a0); hi = f4_add(hi, a1); }
        if (r1) { bf8_unpack(__ldg(r1 + off), a0, a1); lo = f4_add(lo, a0); hi = f4_add(hi, a1); }
        if (r2) { bf8_unpack(__ldg(r2 + off), a0, a1); lo = f4_add(lo, a0); hi = f4_add(hi, a1); }
        out[off] = bf8_pack(lo, hi);
    }
}

extern "C" int paif_dilconv_forward_bf16(const void* x, const float* dw, const float* pw, const float* ch_scale,
                                         const float* ch_shift, const void* r1, const void* r2, void* out, int add_x,
                                         int C, int k, int dil, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(x && dw && pw && out, "null pointer");
    PAIF_REQUIRE(C == 32, "C must be 32");
    PAIF_REQUIRE(B > 0 && B <= 65535, "B out of range");
    PAIF_REQUIRE((long long)H * W < (1ll << 31), "image too large");
#define DC_CASE(K_, D_)                                                                                              \
    if (k == K_ && dil == D_) {                                                                                      \
        dilconv_fused_bf16_kernel<K_, D_><<<pix_grid(W, H, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(              \
            static_cast<const uint4*>(x), dw, pw, ch_scale, ch_shift, static_cast<const uint4*>(r1),                 \
            static_cast<const uint4*>(r2), static_cast<uint4*>(out), add_x, H, W);                                   \
        return check_launch("paif_dilconv_forward_bf16");                                                            \
    }
    DC_CASE(3, 1) DC_CASE(3, 2)
#undef DC_CASE
    set_error("paif_dilconv_forward_bf16: kernel %d dilation %d not instantiated (3x3, dilation 1 or 2)", k, dil);
    return PAIF_ENOTSUP;
}

extern "C" int paif_dwconv_forward(const float* x, const float* w, int relu_in, const float* mask_src,
                                   const float* post_res, float* out, int C, int k, int dil,
                                   int B, int H, int W, void* stream) {
    PAIF_REQUIRE(x && w && out, "null pointer");
    PAIF_REQUIRE(C % 4 == 0 && k >= 1 && k <= 7 && (k & 1), "unsupported C / kernel size");
    PAIF_REQUIRE((long long)B * (C / 4) <= 65535, "B*C/4 exceeds grid.z");
    dwconv_kernel<<<pix_grid(W, H, B * (C / 4)), dim3(32, 8), 0, ST>>>(x, w, relu_in, mask_src, post_res, out,
                                                                    C / 4, k, dil, H, W);
    return check_launch("paif_dwconv_forward");
}

namespace paif {
int dilconv_tc_launch(const float* x, const float* dw, const float* pw, const float* ch_scale, const float* ch_shift,
                      const float* r1, const float* r2, float* out, int add_x, int dil, int B, int H, int W,
                      cudaStream_t stream);
}

extern "C" int paif_dilconv_forward(const float* x, const float* dw, const float* pw, const float* ch_scale,
                                    const float* ch_shift, const float* r1, const float* r2, float* out,
                                    int add_x, int engine, int C, int k, int dil, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(x && dw && pw && out, "null pointer");
    PAIF_REQUIRE(C == 32, "C must be 32");
    PAIF_REQUIRE(k >= 1 && k <= 7 && (k & 1) && dil >= 1, "bad kernel size / dilation");
    PAIF_REQUIRE(B > 0 && B <= 65535, "B out of range");
    PAIF_REQUIRE((long long)H * W < (1ll << 31), "image too large");
    // the tensor-core variant is correct but measured slower (1.00 vs 0.90 ms at 16x480x640: the depthwise taps'
    // L1 traffic dominates either way), so AUTO keeps the FFMA kernel; PAIF_ENGINE_TCGEN05 selects it explicitly
    if (engine == PAIF_ENGINE_TCGEN05 && k == 3 && (dil == 1 || dil == 2))
        return dilconv_tc_launch(x, dw, pw, ch_scale, ch_shift, r1, r2, out, add_x, dil, B, H, W, ST);
#define DC_CASE(K_, D_)                                                                                              \
    if (k == K_ && dil == D_) {                                                                                      \
        dilconv_fused_kernel<K_, D_><<<pix_grid(W, H, B), dim3(32, 8), 0, ST>>>(x, dw, pw, ch_scale, ch_shift, r1, r2, out, add_x, H, W); \
        return check_launch("paif_dilconv_forward");                                                                 \
    }
    DC_CASE(3, 1) DC_CASE(3, 2)
#undef DC_CASE
    set_error("paif_dilconv_forward: kernel %d dilation %d not instantiated (3x3, dilation 1 or 2)", k, dil);
    return PAIF_ENOTSUP;
}

// ------------------------------------------------------------------------------------------
// ChannelPool + spatial attention + blend in ONE kernel (core/model_fusion_auto.py:1352-1368, :631-632):
// a 32 x 16 output tile computes the 4 pooled planes on its halo'd region into shared memory, then the k x k
// conv, the sigmoid and the blend.  Both maps are read once from HBM (halo re-reads come from L2) and the
// pooled plane never leaves the SM: 3 map transfers instead of 5.
// ------------------------------------------------------------------------------------------
constexpr int SF_TX = 32, SF_TY = 16;

template <int QC>                                                // QC = C / 4 when known at compile time (8), else 0
__global__ void __launch_bounds__(256)
spa_fused_kernel(const float* __restrict__ w, int k, const float* __restrict__ a, const float* __restrict__ v,
                 float* __restrict__ agg, float* __restrict__ scale_out, int Qrt, int H, int W) {
    const int Q = QC ? QC : Qrt;
    __shared__ float4 sw[49];
    __shared__ float4 sp[(SF_TY + 6) * (SF_TX + 6)];            // pooled (max_a, mean_a, max_v, mean_v), halo <= 3
    const int tid = threadIdx.x;
    const int taps = k * k, pad = (k - 1) / 2;
    if (tid < taps) sw[tid] = make_float4(w[tid], w[taps + tid], w[2 * taps + tid], w[3 * taps + tid]);
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * SF_TX, y0 = blockIdx.y * SF_TY;
    const int RX = SF_TX + 2 * pad, RY = SF_TY + 2 * pad;
    const size_t plane = (size_t)H * W;
    const float4* ap = reinterpret_cast<const float4*>(a) + (size_t)b * Q * plane;
    const float4* vp = reinterpret_cast<const float4*>(v) + (size_t)b * Q * plane;
    const float inv = 1.f / (float)(Q * 4);
    for (int i = tid; i < RX * RY; i += 256) {
        const int ry = i / RX, rx = i - ry * RX;
        const int yy = y0 - pad + ry, xx = x0 - pad + rx;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);             // zero padding of the conv input
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const size_t pix = (size_t)yy * W + xx;
            float amax = -INFINITY, asum = 0.f, vmax = -INFINITY, vsum = 0.f;
#pragma unroll
            for (int q = 0; q < (QC ? QC : 1); ++q) {
                for (int qq = q; qq < Q; qq += (QC ? Q : 1)) {          // compile-time trip count when QC != 0
                    const float4 t = __ldg(ap + qq * plane + pix);
                    amax = fmaxf(fmaxf(amax, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
                    asum += (t.x + t.y) + (t.z + t.w);
                    const float4 u = __ldg(vp + qq * plane + pix);
                    vmax = fmaxf(fmaxf(vmax, fmaxf(u.x, u.y)), fmaxf(u.z, u.w));
                    vsum += (u.x + u.y) + (u.z + u.w);
                }
            }
            p = make_float4(amax, asum * inv, vmax, vsum * inv);
        }
        sp[i] = p;
    }
    __syncthreads();
    for (int o = tid; o < SF_TX * SF_TY; o += 256) {
        const int oy = o / SF_TX, ox = o - oy * SF_TX;
        const int y = y0 + oy, x = x0 + ox;
        if (y >= H || x >= W) continue;
        float acc = 0.f;
        for (int ty = 0; ty < k; ++ty)
            for (int tx = 0; tx < k; ++tx) {
                const float4 p = sp[(oy + ty) * RX + ox + tx];
                const float4 ww = sw[ty * k + tx];
                acc = fmaf(p.x, ww.x, acc); acc = fmaf(p.y, ww.y, acc);
                acc = fmaf(p.z, ww.z, acc); acc = fmaf(p.w, ww.w, acc);
            }
        const float s = sigmoid_f(acc), s1 = 1.f - s;
        const size_t pix = (size_t)y * W + x;
        if (scale_out) scale_out[(size_t)b * plane + pix] = s;
        float4* op = reinterpret_cast<float4*>(agg) + (size_t)b * Q * plane + pix;
#pragma unroll
        for (int q = 0; q < (QC ? QC : 1); ++q) {
            for (int qq = q; qq < Q; qq += (QC ? Q : 1)) {
                const float4 t = __ldg(ap + qq * plane + pix), u = __ldg(vp + qq * plane + pix);
                op[qq * plane] = make_float4(s * t.x + s1 * u.x, s * t.y + s1 * u.y, s * t.z + s1 * u.z, s * t.w + s1 * u.w);
            }
        }
    }
}

// C = 32 version that reads both maps ONCE from HBM: a thread keeps its own pixel's 2 x 32 values in registers between
// the pooling phase and the blend (ncu on the two-pass version above at 16x480x640: 2.70 GB of DRAM reads for 1.26 GB
// of maps — with ~150 MB of tiles in flight between the phases the re-read misses the 126 MB L2 — at 6.5 TB/s, i.e.
// HBM-bound on avoidable traffic).  Tile 32 x 8, one pixel per thread; halo pixels are pooled first (temporaries), the
// own pixel last, so only 64 data registers are live across the barrier.  BF: bf16 C8 maps.
constexpr int SR_TX = 32, SR_TY = 8;

__device__ __forceinline__ float4 pool_px32(const float4 (&t)[8], const float4 (&u)[8]) {
    float amax = -INFINITY, asum = 0.f, vmax = -INFINITY, vsum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        amax = fmaxf(fmaxf(amax, fmaxf(t[q].x, t[q].y)), fmaxf(t[q].z, t[q].w));
        asum += (t[q].x + t[q].y) + (t[q].z + t[q].w);
        vmax = fmaxf(fmaxf(vmax, fmaxf(u[q].x, u[q].y)), fmaxf(u[q].z, u[q].w));
        vsum += (u[q].x + u[q].y) + (u[q].z + u[q].w);
    }
    return make_float4(amax, asum * (1.f / 32.f), vmax, vsum * (1.f / 32.f));
}

template <bool BF>
__global__ void __launch_bounds__(256)
spa_fused32_kernel(const float* __restrict__ w, int k, const void* __restrict__ a, const void* __restrict__ v,
                   void* __restrict__ agg, float* __restrict__ scale_out, int H, int W) {
    __shared__ float4 sw[49];
    __shared__ float4 sp[(SR_TY + 6) * (SR_TX + 6)];            // pooled (max_a, mean_a, max_v, mean_v), halo <= 3
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int taps = k * k, pad = (k - 1) / 2;
    if (tid < taps) sw[tid] = make_float4(w[tid], w[taps + tid], w[2 * taps + tid], w[3 * taps + tid]);
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * SR_TX, y0 = blockIdx.y * SR_TY;
    const int RX = SR_TX + 2 * pad, RY = SR_TY + 2 * pad;
    const size_t plane = (size_t)H * W;
    const void* ap = img32<BF>(a, b, plane);
    const void* vp = img32<BF>(v, b, plane);
    // halo ring of the tile: pad full-width rows above and below, 2*pad columns beside each tile row
    const int nh = RX * RY - SR_TX * SR_TY, band = pad * RX;
    for (int i = tid; i < nh; i += 256) {
        int ry, rx;
        if (i < band) { ry = i / RX; rx = i - ry * RX; }
        else if (i < 2 * band) { const int j = i - band; ry = j / RX; rx = j - ry * RX; ry += pad + SR_TY; }
        else { const int j = i - 2 * band; ry = j / (2 * pad); const int c = j - ry * 2 * pad; ry += pad; rx = c < pad ? c : c + SR_TX; }
        const int yy = y0 - pad + ry, xx = x0 - pad + rx;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);             // zero padding of the conv input
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            float4 ht[8], hu[8];
            ld_px32<BF>(ap, plane, (size_t)yy * W + xx, ht);
            ld_px32<BF>(vp, plane, (size_t)yy * W + xx, hu);
            p = pool_px32(ht, hu);
        }
        sp[ry * RX + rx] = p;
    }
    const int x = x0 + tx, y = y0 + ty;
    const bool in = x < W && y < H;
    const size_t pix = (size_t)y * W + x;
    float4 t[8], u[8];
    float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) {
        ld_px32<BF>(ap, plane, pix, t);
        ld_px32<BF>(vp, plane, pix, u);
        own = pool_px32(t, u);
    }
    sp[(ty + pad) * RX + tx + pad] = own;
    __syncthreads();
    if (!in) return;
    float acc = 0.f;
    for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
            const float4 p = sp[(ty + dy) * RX + tx + dx];
            const float4 ww = sw[dy * k + dx];
            acc = fmaf(p.x, ww.x, acc); acc = fmaf(p.y, ww.y, acc);
            acc = fmaf(p.z, ww.z, acc); acc = fmaf(p.w, ww.w, acc);
        }
    const float s = sigmoid_f(acc), s1 = 1.f - s;
    if (scale_out) scale_out[(size_t)b * plane + pix] = s;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        t[q] = make_float4(s * t[q].x + s1 * u[q].x, s * t[q].y + s1 * u[q].y, s * t[q].z + s1 * u[q].z, s * t[q].w + s1 * u[q].w);
    st_px32<BF>(img32<BF>(agg, b, plane), plane, pix, t);
}

// bf16 C8 maps (C = 32): the two-pass structure is kept — half the bytes are in flight between the phases, so the re-read
// hits L2 (measured: 0.215 ms vs 0.267 ms for the register-retaining kernel above instantiated for bf16)
__global__ void __launch_bounds__(256)
spa_fused_bf16_kernel(const float* __restrict__ w, int k, const void* __restrict__ a, const void* __restrict__ v,
                      void* __restrict__ agg, int H, int W) {
    __shared__ float4 sw[49];
    __shared__ float4 sp[(SF_TY + 6) * (SF_TX + 6)];
    const int tid = threadIdx.x;
    const int taps = k * k, pad = (k - 1) / 2;
    if (tid < taps) sw[tid] = make_float4(w[tid], w[taps + tid], w[2 * taps + tid], w[3 * taps + tid]);
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * SF_TX, y0 = blockIdx.y * SF_TY;
    const int RX = SF_TX + 2 * pad, RY = SF_TY + 2 * pad;
    const size_t plane = (size_t)H * W;
    const void* ap = img32<true>(a, b, plane);
    const void* vp = img32<true>(v, b, plane);
    for (int i = tid; i < RX * RY; i += 256) {
        const int ry = i / RX, rx = i - ry * RX;
        const int yy = y0 - pad + ry, xx = x0 - pad + rx;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const size_t pix = (size_t)yy * W + xx;
            float4 t[8], u[8];
            ld_px32<true>(ap, plane, pix, t);
            ld_px32<true>(vp, plane, pix, u);
            float amax = -INFINITY, asum = 0.f, vmax = -INFINITY, vsum = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                amax = fmaxf(fmaxf(amax, fmaxf(t[q].x, t[q].y)), fmaxf(t[q].z, t[q].w));
                asum += (t[q].x + t[q].y) + (t[q].z + t[q].w);
                vmax = fmaxf(fmaxf(vmax, fmaxf(u[q].x, u[q].y)), fmaxf(u[q].z, u[q].w));
                vsum += (u[q].x + u[q].y) + (u[q].z + u[q].w);
            }
            p = make_float4(amax, asum * (1.f / 32.f), vmax, vsum * (1.f / 32.f));
        }
        sp[i] = p;
    }
    __syncthreads();
    for (int o = tid; o < SF_TX * SF_TY; o += 256) {
        const int oy = o / SF_TX, ox = o - oy * SF_TX;
        const int y = y0 + oy, x = x0 + ox;
        if (y >= H || x >= W) continue;
        float acc = 0.f;
        for (int ty = 0; ty < k; ++ty)
            for (int tx = 0; tx < k; ++tx) {
                const float4 p = sp[(oy + ty) * RX + ox + tx];
                const float4 ww = sw[ty * k + tx];
                acc = fmaf(p.x, ww.x, acc); acc = fmaf(p.y, ww.y, acc);
                acc = fmaf(p.z, ww.z, acc); acc = fmaf(p.w, ww.w, acc);
            }
        const float s = sigmoid_f(acc), s1 = 1.f - s;
        const size_t pix = (size_t)y * W + x;
        float4 t[8], u[8];
        ld_px32<true>(ap, plane, pix, t);
        ld_px32<true>(vp, plane, pix, u);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            t[q] = make_float4(s * t[q].x + s1 * u[q].x, s * t[q].y + s1 * u[q].y, s * t[q].z + s1 * u[q].z, s * t[q].w + s1 * u[q].w);
        st_px32<true>(img32<true>(agg, b, plane), plane, pix, t);
    }
}

extern "C" int paif_spa_fused_forward_bf16(const float* w, int k, const void* ir_f, const void* vis_f, void* agg,
                                           int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(w && ir_f && vis_f && agg, "null pointer");
    PAIF_REQUIRE(C == 32 && k >= 1 && k <= 7 && (k & 1), "unsupported C / kernel size");
    PAIF_REQUIRE(B > 0 && B <= 65535, "B out of range");
    spa_fused_bf16_kernel<<<dim3(cdiv(W, SF_TX), cdiv(H, SF_TY), B), 256, 0, ST>>>(w, k, ir_f, vis_f, agg, H, W);
    return check_launch("paif_spa_fused_forward_bf16");
}

extern "C" int paif_spa_fused_forward(const float* w, int k, const float* ir_f, const float* vis_f, float* agg,
                                      float* scale_out, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(w && ir_f && vis_f && agg, "null pointer");
    PAIF_REQUIRE(C % 4 == 0 && k >= 1 && k <= 7 && (k & 1), "unsupported C / kernel size");
    PAIF_REQUIRE(B > 0 && B <= 65535, "B out of range");
    const dim3 grid(cdiv(W, SF_TX), cdiv(H, SF_TY), B);
    if (C == 32)
        spa_fused32_kernel<false><<<dim3(cdiv(W, SR_TX), cdiv(H, SR_TY), B), 256, 0, ST>>>(w, k, ir_f, vis_f, agg, scale_out, H, W);
    else spa_fused_kernel<0><<<grid, 256, 0, ST>>>(w, k, ir_f, vis_f, agg, scale_out, C / 4, H, W);
    return check_launch("paif_spa_fused_forward");
}

extern "C" int paif_channel_pool(const float* ir_f, const float* vis_f, float* pooled,
                                 int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(ir_f && vis_f && pooled, "null pointer");
    PAIF_REQUIRE(C % 4 == 0, "C must be a multiple of 4");
    channel_pool_kernel<<<pix_grid(W, H, B), dim3(32, 8), 0, ST>>>(ir_f, vis_f, pooled, C / 4, H, W);
    return check_launch("paif_channel_pool");
}

extern "C" int paif_spa_blend_forward(const float* pooled, const float* w, int k, const float* ir_f,
                                      const float* vis_f, float* agg, float* scale_out,
                                      int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(pooled && w && ir_f && vis_f && agg, "null pointer");
    PAIF_REQUIRE(C % 4 == 0 && k >= 1 && k <= 7 && (k & 1), "unsupported C / kernel size");
    spa_blend_kernel<<<pix_grid(W, H, B), dim3(32, 8), 0, ST>>>(pooled, w, k, ir_f, vis_f, agg, scale_out, C / 4, H, W);
    return check_launch("paif_spa_blend_forward");
}

extern "C" int paif_spa_blend_backward_pre(const float* gagg, const float* ir_f, const float* vis_f,
                                           const float* scale, float* gpre, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(gagg && ir_f && vis_f && scale && gpre, "null pointer");
    spa_blend_bwd_pre_kernel<<<pix_grid(W, H, B), dim3(32, 8), 0, ST>>>(gagg, ir_f, vis_f, scale, gpre, C / 4, H, W);
    return check_launch("paif_spa_blend_backward_pre");
}

extern "C" int paif_spa_blend_backward(const float* gagg, const float* ir_f, const float* vis_f,
                                       const float* scale, const float* gpre, const float* w, int k,
                                       float* g_ir_f, float* g_vis_f, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(gagg && ir_f && vis_f && scale && gpre && w && g_ir_f && g_vis_f, "null pointer");
    PAIF_REQUIRE(k >= 1 && k <= 7 && (k & 1), "unsupported kernel size");
    spa_blend_bwd_kernel<<<pix_grid(W, H, B), dim3(32, 8), 0, ST>>>(gagg, ir_f, vis_f, scale, gpre, w, k,
                                                                  g_ir_f, g_vis_f, C / 4, H, W);
    return check_launch("paif_spa_blend_backward");
}

extern "C" int paif_eca_scale(const float* chan_partials, int tiles, const float* w1d, int k, float* e,
                              int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(chan_partials && w1d && e, "null pointer");
    PAIF_REQUIRE(C <= 1024 && tiles > 0, "bad C / tiles");
    eca_scale_kernel<<<B, C, C * sizeof(float), ST>>>(chan_partials, tiles, w1d, k, e, C, 1.f / ((float)H * (float)W));
    return check_launch("paif_eca_scale");
}

extern "C" int paif_eca_apply(const float* o, const float* x, const float* e, const float* slope,
                              const float* post_res, float* out, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(o && x && e && slope && out, "null pointer");
    eca_apply_kernel<<<pix_grid(W, H, B * (C / 4)), dim3(32, 8), 0, ST>>>(o, x, e, slope, post_res, out, C / 4, H, W);
    return check_launch("paif_eca_apply");
}

extern "C" int paif_eca_apply_bf16(const void* o, const void* x, const float* e, const float* slope,
                                   const void* post_res, void* out, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(o && x && e && slope && out, "null pointer");
    PAIF_REQUIRE(C % 8 == 0 && (long long)B * (C / 8) <= 65535, "C must be a multiple of 8; B*C/8 <= 65535");
    eca_apply_bf16_kernel<<<pix_grid(W, H, B * (C / 8)), dim3(32, 8), 0, ST>>>(
        static_cast<const uint4*>(o), static_cast<const uint4*>(x), e, slope, static_cast<const uint4*>(post_res),
        static_cast<uint4*>(out), C / 8, H, W);
    return check_launch("paif_eca_apply_bf16");
}

extern "C" int paif_eca_bwd_tiles(int H, int W) { return cdiv(W, 32) * cdiv(H, ECAB_ROWS); }

extern "C" int paif_eca_bwd_pass1(const float* gu, const float* o, const float* x, const float* e,
                                  const float* slope, float* gw, float* partials,
                                  int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(gu && o && x && e && slope && gw && partials, "null pointer");
    PAIF_REQUIRE(C == 32, "C must be 32");
    dim3 grid(cdiv(W, 32), cdiv(H, ECAB_ROWS), B);
    eca_bwd_pass1_kernel<<<grid, dim3(32, 8), 0, ST>>>(gu, o, x, e, slope, gw, partials, C / 4, H, W);
    return check_launch("paif_eca_bwd_pass1");
}

extern "C" int paif_eca_bwd_scale(const float* partials, int tiles, const float* e, const float* w1d, int k,
                                  float* gm, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(partials && e && w1d && gm, "null pointer");
    eca_bwd_scale_kernel<<<B, C, C * sizeof(float), ST>>>(partials, tiles, e, w1d, k, gm, C, 1.f / ((float)H * (float)W));
    return check_launch("paif_eca_bwd_scale");
}

extern "C" int paif_eca_bwd_pass2(const float* gw, const float* e, const float* gm, float* go,
                                  int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(gw && e && gm && go, "null pointer");
    eca_bwd_pass2_kernel<<<pix_grid(W, H, B * (C / 4)), dim3(32, 8), 0, ST>>>(gw, e, gm, go, C / 4, H, W);
    return check_launch("paif_eca_bwd_pass2");
}

static int out_smem_attr() {
    static unsigned long long done = 0;
    int dev;
    if (!attr_needed(done, &dev)) return 0;
    const int bytes = 9 * 25 * OUT_C * sizeof(float);
    cudaError_t e1 = cudaFuncSetAttribute(out_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, OF_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(out_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(out_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OF_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("out kernel smem attr failed"); return (int)(e1 ? e1 : e2); }
    attr_mark(done, dev);
    return 0;
}

extern "C" int paif_out_forward(const float* feat, const float* wm, const float* slope, float* out, float* pre_out,
                                int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && wm && slope && out, "null pointer");
    PAIF_REQUIRE(C == OUT_C, "C must be 32");
    PAIF_REQUIRE(H >= 2 && W >= 2, "H, W must be >= 2");
    if (int r = out_smem_attr()) return r;
    out_forward_kernel<false><<<dim3(cdiv(W, OF_TX), cdiv(H, OF_TY), B), 256, OF_SMEM, ST>>>(feat, wm, slope, out, pre_out, H, W);
    if (int r = check_launch("paif_out_forward")) return r;
    out_border_kernel<false><<<dim3(cdiv(2 * W + 2 * (H - 2), 128), B), 128, 0, ST>>>(feat, wm, slope, out, pre_out, H, W);
    return check_launch("paif_out_forward");
}

namespace paif {
int out_tc_launch(const void* feat, const void* wmma, const float* slope, float* out, float* pre_out,
                  int bf16, int B, int H, int W, cudaStream_t stream);
}

extern "C" int paif_out_forward_tc(const void* feat, const void* w_mma, const float* wm, const float* slope,
                                   float* out, float* pre_out, int storage, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && w_mma && wm && slope && out, "null pointer");
    PAIF_REQUIRE(C == OUT_C, "C must be 32");
    PAIF_REQUIRE(H >= 2 && W >= 2, "H, W must be >= 2");
    PAIF_REQUIRE(storage == PAIF_STORAGE_F32 || storage == PAIF_STORAGE_BF16, "storage must be F32 or BF16");
    PAIF_REQUIRE(B > 0 && B <= 65535, "B out of range");
    if (int r = out_tc_launch(feat, w_mma, slope, out, pre_out, storage == PAIF_STORAGE_BF16, B, H, W, ST)) return r;
    const dim3 bgrid(cdiv(2 * W + 2 * (H - 2), 128), B);
    if (storage == PAIF_STORAGE_BF16) out_border_kernel<true><<<bgrid, 128, 0, ST>>>(feat, wm, slope, out, pre_out, H, W);
    else out_border_kernel<false><<<bgrid, 128, 0, ST>>>(feat, wm, slope, out, pre_out, H, W);
    return check_launch("paif_out_forward_tc");
}

extern "C" int paif_out_forward_bf16(const void* feat, const float* wm, const float* slope, float* out,
                                     int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && wm && slope && out, "null pointer");
    PAIF_REQUIRE(C == OUT_C, "C must be 32");
    PAIF_REQUIRE(H >= 2 && W >= 2, "H, W must be >= 2");
    if (int r = out_smem_attr()) return r;
    out_forward_kernel<true><<<dim3(cdiv(W, OF_TX), cdiv(H, OF_TY), B), 256, OF_SMEM, ST>>>(feat, wm, slope, out, nullptr, H, W);
    if (int r = check_launch("paif_out_forward_bf16")) return r;
    out_border_kernel<true><<<dim3(cdiv(2 * W + 2 * (H - 2), 128), B), 128, 0, ST>>>(feat, wm, slope, out, nullptr, H, W);
    return check_launch("paif_out_forward_bf16");
}

extern "C" int paif_out_backward(const float* g, const float* out, const float* pre_out, const float* wm,
                                 const float* slope, float* gfeat, const float* mask_src, const float* mask_slope,
                                 float* gfeat_masked, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(g && out && pre_out && wm && slope && gfeat, "null pointer");
    PAIF_REQUIRE(C == OUT_C, "C must be 32");
    PAIF_REQUIRE(!gfeat_masked || mask_src, "gfeat_masked needs mask_src");
    if (int r = out_smem_attr()) return r;
    out_backward_kernel<<<pix_grid(W, H, B), dim3(32, 8), 9 * 25 * OUT_C * sizeof(float), ST>>>(
        g, out, pre_out, wm, slope, gfeat, mask_src, mask_slope, gfeat_masked, H, W);
    return check_launch("paif_out_backward");
}

extern "C" int paif_mask_scale(const float* g, const float* mask_src, const float* mask_slope, float scale,
                               float* out, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(g && out, "null pointer");
    const size_t n4 = (size_t)B * (C / 4) * H * W;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    mask_scale_kernel<<<blocks, 256, 0, ST>>>(g, mask_src, mask_slope, scale, out, n4);
    return check_launch("paif_mask_scale");
}

__global__ void __launch_bounds__(256)
add_act_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ slope_p,
               float* __restrict__ out, float* __restrict__ pre_out, size_t n4) {
    const float a = *slope_p;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 t = f4_add(reinterpret_cast<const float4*>(x)[i], reinterpret_cast<const float4*>(y)[i]);
        if (pre_out) reinterpret_cast<float4*>(pre_out)[i] = t;
        reinterpret_cast<float4*>(out)[i] = make_float4(prelu_f(t.x, a), prelu_f(t.y, a), prelu_f(t.z, a), prelu_f(t.w, a));
    }
}

extern "C" int paif_add_act(const float* x, const float* y, const float* slope, float* out, float* pre_out,
                            long long n, void* stream) {
    PAIF_REQUIRE(x && y && slope && out, "null pointer");
    PAIF_REQUIRE(n >= 0 && n % 4 == 0, "n must be a multiple of 4");
    if (n == 0) return 0;
    const size_t n4 = (size_t)n / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    add_act_kernel<<<blocks, 256, 0, ST>>>(x, y, slope, out, pre_out, n4);
    return check_launch("paif_add_act");
}

// bf16 C8 map [B][C/8][H][W][8] -> fp32 C4 map [B][C/4][H][W][4]: how the backward-to-input chain (fp32 gradient maps)
// reads the activations a bf16-storage forward saved.  One thread = one pixel of one oct.
__global__ void __launch_bounds__(256)
widen_bf16_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, long long plane, long long total) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long oct = i / plane, pix = i - oct * plane;          // oct counts (image, oct) pairs
        float4 lo, hi;
        bf8_unpack(__ldg(src + i), lo, hi);
        dst[(2 * oct) * plane + pix] = lo;
        dst[(2 * oct + 1) * plane + pix] = hi;
    }
}

extern "C" int paif_widen_bf16_map(const void* src, float* dst, int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(src && dst, "null pointer");
    PAIF_REQUIRE(C > 0 && C % 8 == 0 && B > 0 && H > 0 && W > 0, "bad shape");
    const long long plane = (long long)H * W, total = plane * (C / 8) * B;
    long long want = (total + 255) / 256;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    widen_bf16_kernel<<<blocks, 256, 0, ST>>>(static_cast<const uint4*>(src), reinterpret_cast<float4*>(dst), plane, total);
    return check_launch("paif_widen_bf16_map");
}

extern "C" int paif_add_maps(const float* a, const float* b, const float* c, float* out, long long n, void* stream) {
    PAIF_REQUIRE(a && b && out, "null pointer");
    PAIF_REQUIRE(n >= 0 && n % 4 == 0, "n must be a multiple of 4");
    if (n == 0) return 0;
    const size_t n4 = (size_t)n / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    add_maps_kernel<<<blocks, 256, 0, ST>>>(a, b, c, out, n4);
    return check_launch("paif_add_maps");
}

extern "C" int paif_confusion_accumulate(const long long* label, const long long* pred, long long count,
                                         int num_classes, long long* conf, void* stream) {
    PAIF_REQUIRE(label && pred && conf, "null pointer");
    PAIF_REQUIRE(num_classes > 0 && num_classes <= 64 && count >= 0, "bad num_classes / count");
    if (count == 0) return 0;
    long long want = (count + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    confusion_kernel<<<blocks, 256, num_classes * num_classes * sizeof(unsigned int), ST>>>(
        label, pred, count, num_classes, reinterpret_cast<unsigned long long*>(conf));
    return check_launch("paif_confusion_accumulate");
}
