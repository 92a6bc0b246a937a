// Kernels either side of the fusion hot path: the PGD step of attack/attack.py:504-512.
#include "common.cuh"

namespace paif {

// delta <- clamp(clamp(delta + alpha * sign(grad), -eps, eps), 0 - x, 1 - x), in place, for one image tensor.
// attack/attack.py:504-512 does this with six elementwise launches per modality (sign, mul-add, three clamps, the
// .data assignment); `grad` is delta.grad, which the reference never zeroes (it steps along the sign of the RUNNING
// SUM of gradients) — the accumulation stays with autograd, this kernel only reads it.
// torch.sign(0) == 0 and torch.clamp(x, min=lo, max=hi) == min(max(x, lo), hi) are followed exactly.
__global__ void __launch_bounds__(256)
pgd_step_kernel(float* __restrict__ delta, const float* __restrict__ grad, const float* __restrict__ x,
                float alpha, float eps, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            float4 d = ld4(delta + i);
            const float4 g = ld4(grad + i), xv = ld4(x + i);
            float* dp = &d.x;
            const float* gp = &g.x;
            const float* xp = &xv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float s = gp[k] > 0.f ? 1.f : (gp[k] < 0.f ? -1.f : 0.f);
                float v = __fadd_rn(dp[k], __fmul_rn(alpha, s));
                v = fminf(fmaxf(v, -eps), eps);
                dp[k] = fminf(fmaxf(v, __fsub_rn(0.f, xp[k])), __fsub_rn(1.f, xp[k]));
            }
            st4(delta + i, d);
        } else {
            for (long long j = i; j < n; ++j) {
                const float gj = grad[j];
                const float s = gj > 0.f ? 1.f : (gj < 0.f ? -1.f : 0.f);
                float v = __fadd_rn(delta[j], __fmul_rn(alpha, s));
                v = fminf(fmaxf(v, -eps), eps);
                delta[j] = fminf(fmaxf(v, __fsub_rn(0.f, x[j])), __fsub_rn(1.f, x[j]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Colour / normalisation glue of the task wrappers (core/model_fusion_auto.py:69-111 RGB2YCrCb / YCrCb2RGB, :712-728
// Network_MM_CompModel.forward): the ~15 eager kernels between the fusion net and the segmentation consumer as
// two passes (min-max partials, then normalise), and their adjoint as two passes.
//   Y = .299 R + .587 G + .114 B; Cr = (R - Y) .713 + .5; Cb = (B - Y) .564 + .5          (visible image)
//   rgb = [fused, Cr - .5, Cb - .5] . [[1,1,1],[1.403,-.714,0],[0,-.344,1.773]]; r = clamp(rgb, 0, 1)
//   t = (r - lo) / (hi - lo) with lo / hi the min / max over the sample (per_sample) or over the whole batch
//   x_c = (255 t_c - mean_c) / std_c
// ------------------------------------------------------------------------------------------------------------
constexpr int GLUE_NT = 256;

struct GlueRaw { float v[3]; };
__device__ __forceinline__ GlueRaw glue_raw(float yf, float R, float G, float B) {
    const float Y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, R), __fmul_rn(0.587f, G)), __fmul_rn(0.114f, B));
    const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(R, Y), 0.713f), 0.5f), -0.5f);
    const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(B, Y), 0.564f), 0.5f), -0.5f);
    GlueRaw o;
    o.v[0] = fmaf(c1, 1.403f, yf);
    o.v[1] = fmaf(c2, -0.344f, fmaf(c1, -0.714f, yf));
    o.v[2] = fmaf(c2, 1.773f, yf);
    return o;
}
__device__ __forceinline__ float clamp01(float v) { return v > 1.f ? 1.f : (v < 0.f ? 0.f : v); }

// fixed-order block reductions (deterministic): shuffle tree inside a warp, then warp 0 over the warp results
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch /* [GLUE_NT / 32] */) {
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = scratch[0];
    for (int w = 1; w < GLUE_NT / 32; ++w) r = op(r, scratch[w]);
    return r;
}

// pass 1: per-block min / max of the clamped RGB image.  partial: [B][nblk][2]
__global__ void __launch_bounds__(GLUE_NT)
glue_minmax_kernel(const float* __restrict__ fused, const float* __restrict__ vis, float* __restrict__ partial, int HW) {
    __shared__ float sc[GLUE_NT / 32];
    const int b = blockIdx.y;
    const float* f = fused + (size_t)b * HW;
    const float* v = vis + (size_t)b * 3 * HW;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = blockIdx.x * GLUE_NT + threadIdx.x; i < HW; i += gridDim.x * GLUE_NT) {
        const GlueRaw r = glue_raw(f[i], v[i], v[HW + i], v[2 * HW + i]);
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float t = clamp01(r.v[c]); lo = fminf(lo, t); hi = fmaxf(hi, t); }
    }
    lo = block_reduce(lo, [](float a, float c) { return fminf(a, c); }, sc);
    hi = block_reduce(hi, [](float a, float c) { return fmaxf(a, c); }, sc);
    if (threadIdx.x == 0) { partial[((size_t)b * gridDim.x + blockIdx.x) * 2] = lo; partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = hi; }
}

// lo / hi of this block's sample (per_sample) or of the whole batch, from the pass-1 partials (every block reduces them
// redundantly in the same fixed order)
__device__ __forceinline__ void glue_lohi(const float* __restrict__ partial, int nblk, int B, int per_sample, int b,
                                          float* sc, float& lo, float& hi) {
    const int first = per_sample ? b * nblk : 0, count = per_sample ? nblk : B * nblk;
    float l = INFINITY, h = -INFINITY;
    for (int i = threadIdx.x; i < count; i += GLUE_NT) { l = fminf(l, partial[(size_t)(first + i) * 2]); h = fmaxf(h, partial[(size_t)(first + i) * 2 + 1]); }
    lo = block_reduce(l, [](float a, float c) { return fminf(a, c); }, sc);
    hi = block_reduce(h, [](float a, float c) { return fmaxf(a, c); }, sc);
}

// pass 2: x = ((r - lo) / (hi - lo) * 255 - mean) / std; ties: [B][nblk][2] = number of elements equal to lo / hi (for the adjoint)
__global__ void __launch_bounds__(GLUE_NT)
glue_normalise_kernel(const float* __restrict__ fused, const float* __restrict__ vis, const float* __restrict__ partial,
                      float* __restrict__ x, int* __restrict__ ties, float* __restrict__ lohi,
                      float m0, float m1, float m2, float s0, float s1, float s2, int HW, int B, int per_sample) {
    __shared__ float sc[GLUE_NT / 32];
    __shared__ int sci[GLUE_NT / 32];
    const int b = blockIdx.y, nblk = gridDim.x;
    float lo, hi;
    glue_lohi(partial, nblk, B, per_sample, b, sc, lo, hi);
    if (blockIdx.x == 0 && threadIdx.x == 0) { lohi[b * 2] = lo; lohi[b * 2 + 1] = hi; }
    const float den = __fsub_rn(hi, lo);
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    const float* f = fused + (size_t)b * HW;
    const float* v = vis + (size_t)b * 3 * HW;
    float* xo = x + (size_t)b * 3 * HW;
    int nlo = 0, nhi = 0;
    for (int i = blockIdx.x * GLUE_NT + threadIdx.x; i < HW; i += nblk * GLUE_NT) {
        const GlueRaw r = glue_raw(f[i], v[i], v[HW + i], v[2 * HW + i]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float rc = clamp01(r.v[c]);
            nlo += rc == lo; nhi += rc == hi;
            const float t = __fdiv_rn(__fsub_rn(rc, lo), den);
            xo[(size_t)c * HW + i] = __fdiv_rn(__fsub_rn(__fmul_rn(t, 255.f), mean[c]), stdv[c]);
        }
    }
    nlo = block_reduce(nlo, [](int a, int c) { return a + c; }, sci);
    nhi = block_reduce(nhi, [](int a, int c) { return a + c; }, sci);
    if (threadIdx.x == 0) { ties[((size_t)b * nblk + blockIdx.x) * 2] = nlo; ties[((size_t)b * nblk + blockIdx.x) * 2 + 1] = nhi; }
}

// adjoint pass 1: per-block sums of g_t (t - 1) and g_t t (the gradients that reach lo and hi).  sums: [B][nblk][2]
__global__ void __launch_bounds__(GLUE_NT)
glue_bwd_sums_kernel(const float* __restrict__ fused, const float* __restrict__ vis, const float* __restrict__ gx,
                     const float* __restrict__ lohi, float* __restrict__ sums,
                     float s0, float s1, float s2, int HW) {
    __shared__ float sc[GLUE_NT / 32];
    const int b = blockIdx.y;
    const float lo = lohi[b * 2], den = __fsub_rn(lohi[b * 2 + 1], lo);
    const float gs[3] = {255.f / s0, 255.f / s1, 255.f / s2};
    const float* f = fused + (size_t)b * HW;
    const float* v = vis + (size_t)b * 3 * HW;
    const float* g = gx + (size_t)b * 3 * HW;
    float a0 = 0.f, a1 = 0.f;
    for (int i = blockIdx.x * GLUE_NT + threadIdx.x; i < HW; i += gridDim.x * GLUE_NT) {
        const GlueRaw r = glue_raw(f[i], v[i], v[HW + i], v[2 * HW + i]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t = __fdiv_rn(__fsub_rn(clamp01(r.v[c]), lo), den);
            const float gt = g[(size_t)c * HW + i] * gs[c];
            a0 = fmaf(gt, t - 1.f, a0);
            a1 = fmaf(gt, t, a1);
        }
    }
    a0 = block_reduce(a0, [](float a, float c) { return a + c; }, sc);
    a1 = block_reduce(a1, [](float a, float c) { return a + c; }, sc);
    if (threadIdx.x == 0) { sums[((size_t)b * gridDim.x + blockIdx.x) * 2] = a0; sums[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = a1; }
}

// adjoint pass 2: gradients w.r.t. the fused image and (through Cr / Cb) the visible image
__global__ void __launch_bounds__(GLUE_NT)
glue_bwd_kernel(const float* __restrict__ fused, const float* __restrict__ vis, const float* __restrict__ gx,
                const float* __restrict__ lohi, const float* __restrict__ sums, const int* __restrict__ ties,
                float* __restrict__ gfused, float* __restrict__ gvis,
                float s0, float s1, float s2, int HW, int B, int per_sample) {
    __shared__ float sc[GLUE_NT / 32];
    __shared__ int sci[GLUE_NT / 32];
    const int b = blockIdx.y, nblk = gridDim.x;
    const float lo = lohi[b * 2], hi = lohi[b * 2 + 1], den = __fsub_rn(hi, lo);
    // gradients of lo / hi and the number of elements that share them (torch.min / torch.max distribute evenly over ties)
    const int first = per_sample ? b * nblk : 0, count = per_sample ? nblk : B * nblk;
    float a0 = 0.f, a1 = 0.f;
    int n0 = 0, n1 = 0;
    for (int i = threadIdx.x; i < count; i += GLUE_NT) {
        a0 += sums[(size_t)(first + i) * 2]; a1 += sums[(size_t)(first + i) * 2 + 1];
        n0 += ties[(size_t)(first + i) * 2]; n1 += ties[(size_t)(first + i) * 2 + 1];
    }
    a0 = block_reduce(a0, [](float a, float c) { return a + c; }, sc);
    a1 = block_reduce(a1, [](float a, float c) { return a + c; }, sc);
    n0 = block_reduce(n0, [](int a, int c) { return a + c; }, sci);
    n1 = block_reduce(n1, [](int a, int c) { return a + c; }, sci);
    const float glo = (a0 / den) / (float)max(n0, 1), ghi = (-a1 / den) / (float)max(n1, 1);
    const float gs[3] = {255.f / s0, 255.f / s1, 255.f / s2};
    const float* f = fused + (size_t)b * HW;
    const float* v = vis + (size_t)b * 3 * HW;
    const float* g = gx + (size_t)b * 3 * HW;
    for (int i = blockIdx.x * GLUE_NT + threadIdx.x; i < HW; i += nblk * GLUE_NT) {
        const GlueRaw r = glue_raw(f[i], v[i], v[HW + i], v[2 * HW + i]);
        float gr[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float rc = clamp01(r.v[c]);
            float t = g[(size_t)c * HW + i] * gs[c] / den;
            if (rc == lo) t += glo;
            if (rc == hi) t += ghi;
            gr[c] = (r.v[c] > 1.f || r.v[c] < 0.f) ? 0.f : t;             // torch.where(x > 1, ones, x) / (x < 0, zeros, x)
        }
        const float gy = gr[0] + gr[1] + gr[2];
        const float gc1 = 1.403f * gr[0] - 0.714f * gr[1], gc2 = -0.344f * gr[1] + 1.773f * gr[2];
        gfused[(size_t)b * HW + i] = gy;
        // Cr = (R - Y) .713 + .5, Cb = (B - Y) .564 + .5, Y = .299 R + .587 G + .114 B
        const float gyv = -(0.713f * gc1 + 0.564f * gc2);
        float* go = gvis + (size_t)b * 3 * HW;
        go[i] = fmaf(0.299f, gyv, 0.713f * gc1);
        go[HW + i] = 0.587f * gyv;
        go[2 * HW + i] = fmaf(0.114f, gyv, 0.564f * gc2);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Loss head of the attack (attack/attack.py:446-448 + Seg_loss :103-114): bilinear up-sampling of the consumer's
// logits to the label size (F.interpolate, align_corners=False) + cross entropy with ignore_index, and its gradient
// w.r.t. the logits.  One pass over the label pixels computes the per-pixel loss (block partial sums in a fixed
// order) and the gradient of the UP-SAMPLED logits; a second pass gathers it back onto the low-resolution grid —
// every logit sums its <= 9 x 9 contributing label pixels in a fixed order, where the stock bilinear backward
// scatters with float atomics (run-to-run differences in the last bits, which PGD's sign() turns into flipped
// pixels).  K <= 16 classes.
// ------------------------------------------------------------------------------------------------------------
constexpr int SL_MAXK = 16;

__device__ __forceinline__ void bilerp_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;          // area_pixel_compute_source_index, align_corners = false
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

__global__ void __launch_bounds__(GLUE_NT)
segloss_forward_kernel(const float* __restrict__ seg, const long long* __restrict__ label, float* __restrict__ partial,
                       float* __restrict__ gup, int K, int h, int w, int H, int W, long long ignore_index, float inv_norm) {
    __shared__ float sc[GLUE_NT / 32];
    const int b = blockIdx.y;
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const float* sb = seg + (size_t)b * K * h * w;
    float acc = 0.f;
    for (int i = blockIdx.x * GLUE_NT + threadIdx.x; i < H * W; i += gridDim.x * GLUE_NT) {
        const int y = i / W, x = i - y * W;
        int y0, y1, x0, x1;
        float ly, lx;
        bilerp_src(y, sy, h, y0, y1, ly);
        bilerp_src(x, sx, w, x0, x1, lx);
        const float hy = 1.f - ly, hx = 1.f - lx;
        float l[SL_MAXK];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < SL_MAXK; ++k) {
            if (k < K) {
                const float* p = sb + (size_t)k * h * w;
                l[k] = hy * (hx * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * (hx * p[y1 * w + x0] + lx * p[y1 * w + x1]);
                m = fmaxf(m, l[k]);
            }
        }
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < SL_MAXK; ++k) if (k < K) se += expf(l[k] - m);
        const long long lab = label[(size_t)b * H * W + i];
        const bool valid = lab != ignore_index && lab >= 0 && lab < K;
        const float lse = m + logf(se);
        if (valid) {
            float lt = 0.f;
#pragma unroll
            for (int k = 0; k < SL_MAXK; ++k) if (k == (int)lab) lt = l[k];
            acc += lse - lt;
        }
        if (gup) {
#pragma unroll
            for (int k = 0; k < SL_MAXK; ++k)
                if (k < K) gup[((size_t)b * K + k) * H * W + i] = valid ? (expf(l[k] - lse) - (k == (int)lab ? 1.f : 0.f)) * inv_norm : 0.f;
        }
    }
    acc = block_reduce(acc, [](float a, float c) { return a + c; }, sc);
    if (threadIdx.x == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = acc * inv_norm;
}

// gseg[b][k][yl][xl] = gscale * sum over the label pixels whose bilinear footprint contains (yl, xl), fixed order
__global__ void __launch_bounds__(GLUE_NT)
segloss_backward_kernel(const float* __restrict__ gup, const float* __restrict__ gscale, float* __restrict__ gseg,
                        int K, int h, int w, int H, int W, long long total) {
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const float gs = gscale ? *gscale : 1.f;
    const int ry = (H + h - 1) / h + 1, rx = (W + w - 1) / w + 1;           // candidate half-window in label pixels
    for (long long i = (long long)blockIdx.x * GLUE_NT + threadIdx.x; i < total; i += (long long)gridDim.x * GLUE_NT) {
        const int xl = (int)(i % w), yl = (int)((i / w) % h);
        const long long bk = i / ((long long)w * h);
        const float* g = gup + (size_t)bk * H * W;
        const int yc = (int)(((float)yl + 0.5f) / sy), xc = (int)(((float)xl + 0.5f) / sx);
        float acc = 0.f;
        for (int y = max(0, yc - 2 * ry); y <= min(H - 1, yc + 2 * ry); ++y) {
            int y0, y1; float ly;
            bilerp_src(y, sy, h, y0, y1, ly);
            const float wy = (y0 == yl ? 1.f - ly : 0.f) + (y1 == yl ? ly : 0.f);
            if (y0 != yl && y1 != yl) continue;
            for (int x = max(0, xc - 2 * rx); x <= min(W - 1, xc + 2 * rx); ++x) {
                int x0, x1; float lx;
                bilerp_src(x, sx, w, x0, x1, lx);
                const float wx = (x0 == xl ? 1.f - lx : 0.f) + (x1 == xl ? lx : 0.f);
                if (x0 != xl && x1 != xl) continue;
                acc = fmaf(wy * wx, g[(size_t)y * W + x], acc);
            }
        }
        gseg[i] = acc * gs;
    }
}

}  // namespace paif

using namespace paif;

static int glue_blocks(int HW) {
    int n = (HW + GLUE_NT * 4 - 1) / (GLUE_NT * 4);
    return n < 1 ? 1 : (n > 296 ? 296 : n);
}

extern "C" int paif_glue_blocks(int H, int W) { return glue_blocks(H * W); }

extern "C" int paif_glue_forward(const float* fused, const float* vis, const float* mean3, const float* std3,
                                 float* x, float* partial, int* ties, float* lohi, int per_sample,
                                 int B, int H, int W, void* stream) {
    PAIF_REQUIRE(fused && vis && mean3 && std3 && x && partial && ties && lohi, "null pointer");
    PAIF_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "bad shape");
    const int HW = H * W;
    dim3 grid(glue_blocks(HW), B);
    glue_minmax_kernel<<<grid, GLUE_NT, 0, (cudaStream_t)stream>>>(fused, vis, partial, HW);
    if (int r = check_launch("paif_glue_forward(min-max)")) return r;
    glue_normalise_kernel<<<grid, GLUE_NT, 0, (cudaStream_t)stream>>>(fused, vis, partial, x, ties, lohi,
        mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], HW, B, per_sample);
    return check_launch("paif_glue_forward");
}

extern "C" int paif_glue_backward(const float* fused, const float* vis, const float* gx, const float* std3,
                                  const float* lohi, const int* ties, float* sums, float* gfused, float* gvis,
                                  int per_sample, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(fused && vis && gx && std3 && lohi && ties && sums && gfused && gvis, "null pointer");
    PAIF_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "bad shape");
    const int HW = H * W;
    dim3 grid(glue_blocks(HW), B);
    glue_bwd_sums_kernel<<<grid, GLUE_NT, 0, (cudaStream_t)stream>>>(fused, vis, gx, lohi, sums, std3[0], std3[1], std3[2], HW);
    if (int r = check_launch("paif_glue_backward(sums)")) return r;
    glue_bwd_kernel<<<grid, GLUE_NT, 0, (cudaStream_t)stream>>>(fused, vis, gx, lohi, sums, ties, gfused, gvis,
                                                                std3[0], std3[1], std3[2], HW, B, per_sample);
    return check_launch("paif_glue_backward");
}

extern "C" int paif_pgd_step(float* delta, const float* grad, const float* x, float alpha, float eps,
                             long long n, void* stream) {
    PAIF_REQUIRE(delta && grad && x, "null pointer");
    PAIF_REQUIRE(n >= 0, "negative size");
    PAIF_REQUIRE(((reinterpret_cast<uintptr_t>(delta) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(x)) & 15) == 0,
                 "pointers must be 16-byte aligned");
    if (n == 0) return 0;
    long long want = (n / 4 + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
    pgd_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(delta, grad, x, alpha, eps, n);
    return check_launch("paif_pgd_step");
}

extern "C" int paif_segloss_forward(const float* seg, const long long* label, float* partial, float* gup,
                                    long long ignore_index, float inv_norm, int K, int B, int h, int w, int H, int W,
                                    void* stream) {
    PAIF_REQUIRE(seg && label && partial, "null pointer");
    PAIF_REQUIRE(K > 0 && K <= SL_MAXK, "1 <= classes <= 16");
    PAIF_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "bad shape");
    dim3 grid(glue_blocks(H * W), B);
    segloss_forward_kernel<<<grid, GLUE_NT, 0, (cudaStream_t)stream>>>(seg, label, partial, gup, K, h, w, H, W, ignore_index, inv_norm);
    return check_launch("paif_segloss_forward");
}

extern "C" int paif_segloss_backward(const float* gup, const float* gscale, float* gseg,
                                     int K, int B, int h, int w, int H, int W, void* stream) {
    PAIF_REQUIRE(gup && gseg, "null pointer");
    PAIF_REQUIRE(K > 0 && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bad shape");
    const long long total = (long long)B * K * h * w;
    long long want = (total + GLUE_NT - 1) / GLUE_NT;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    segloss_backward_kernel<<<blocks, GLUE_NT, 0, (cudaStream_t)stream>>>(gup, gscale, gseg, K, h, w, H, W, total);
    return check_launch("paif_segloss_backward");
}
