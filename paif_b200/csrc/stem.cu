// stem_1 / stem_2 (Conv2d(1,32,3,pad 1) + PReLU) fused with Cell_Decom.get_residue, and their
// backward-to-input.  Reference: core/model_fusion_auto.py:607-614, :517-521, :628-629.
// HBM-bound: 4 B in, 132 B out per pixel; one thread per pixel, 512-byte coalesced quad stores.
#include "common.cuh"

namespace paif {

constexpr int STEM_C = 32;

// RGB: img points at the R plane and sc is the channel stride; the pixel value is Y = .299 R + .587 G + .114 B
// (RGB2YCrCb, core/model_fusion_auto.py:69-92: separate multiplies and adds, in that order)
template <bool RGB>
__global__ void __launch_bounds__(256)
stem_forward_kernel(const float* __restrict__ img, long long sb, long long sc, long long sy, long long sx,
                    const float* __restrict__ w, const float* __restrict__ slope_p,
                    float* __restrict__ feat, float* __restrict__ residue, uint4* __restrict__ feat16, int H, int W) {
    // weights as channel pairs [pair][tap] = (w[2p][t], w[2p+1][t]): the 288 multiply-adds of a pixel are 144 packed
    // FFMA2 (the kernel is instruction-issue bound; each component is the same fused multiply-add as before)
    __shared__ float2 sw2[(STEM_C / 2) * 9];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < (STEM_C / 2) * 9; i += 256) {
        const int pr = i / 9, t = i - pr * 9;
        sw2[i] = make_float2(w[(2 * pr) * 9 + t], w[(2 * pr + 1) * 9 + t]);
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const float a = *slope_p;
    float in[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const float* px = img + b * sb + yy * sy + xx * sx;
                if (RGB) v = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, px[0]), __fmul_rn(0.587f, px[sc])), __fmul_rn(0.114f, px[2 * sc]));
                else v = px[0];
            }
            in[(dy + 1) * 3 + (dx + 1)] = v;
        }
    float2 in2[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) in2[t] = make_float2(in[t], in[t]);
    float vmax = -INFINITY, vmin = INFINITY;
    const size_t plane = (size_t)H * W;
    float4* out = reinterpret_cast<float4*>(feat) + (size_t)b * (STEM_C / 4) * plane + (size_t)y * W + x;
    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < STEM_C / 4; ++q) {
        float r[4];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
            const float2* wc = sw2 + (q * 2 + jp) * 9;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) acc = __ffma2_rn(in2[t], wc[t], acc);
            const float v0 = prelu_f(acc.x, a), v1 = prelu_f(acc.y, a);
            vmax = fmaxf(vmax, fmaxf(v0, v1));
            vmin = fminf(vmin, fminf(v0, v1));
            r[2 * jp] = v0; r[2 * jp + 1] = v1;
        }
        out[q * plane] = make_float4(r[0], r[1], r[2], r[3]);
        if (feat16) {                 // bf16 C8 copy for the bf16 storage mode (residual of the decomposition branch)
            if (q & 1)
                feat16[((size_t)b * (STEM_C / 8) + (q >> 1)) * plane + (size_t)y * W + x] =
                    bf8_pack(prev, make_float4(r[0], r[1], r[2], r[3]));
            else prev = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
    residue[(size_t)b * plane + (size_t)y * W + x] = vmax - vmin;
}

// total = g0+g1+g2+g3 + routing of the guide gradient to the arg-max / arg-min channel;
// gpre = total * PReLU'(feat)  (sign(feat) == sign(pre-activation) for slope > 0).
__global__ void __launch_bounds__(256)
stem_backward_pre_kernel(const float* __restrict__ feat, const float* __restrict__ slope_p,
                         const float* __restrict__ g0, const float* __restrict__ g1,
                         const float* __restrict__ g2, const float* __restrict__ g3,
                         const float* __restrict__ gres_partial, int nparts, float* __restrict__ gpre,
                         int B, int H, int W) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const float a = *slope_p;
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)y * W + x;
    const size_t base = (size_t)b * (STEM_C / 4) * plane + pix;   // float4 units
    float f[STEM_C];
#pragma unroll
    for (int q = 0; q < STEM_C / 4; ++q) {
        float4 v = reinterpret_cast<const float4*>(feat)[base + q * plane];
        f[q * 4 + 0] = v.x; f[q * 4 + 1] = v.y; f[q * 4 + 2] = v.z; f[q * 4 + 3] = v.w;
    }
    int imax = 0, imin = 0;
    float vmax = f[0], vmin = f[0];
#pragma unroll
    for (int c = 1; c < STEM_C; ++c) {
        if (f[c] > vmax) { vmax = f[c]; imax = c; }
        if (f[c] < vmin) { vmin = f[c]; imin = c; }
    }
    float gres = 0.f;
    if (gres_partial) {
        for (int q = 0; q < nparts; ++q) gres += gres_partial[((size_t)q * B + b) * plane + pix];   // fixed order: deterministic
    }
#pragma unroll
    for (int q = 0; q < STEM_C / 4; ++q) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        const float* gs[4] = {g0, g1, g2, g3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (gs[k]) {
                float4 v = reinterpret_cast<const float4*>(gs[k])[base + q * plane];
                t[0] += v.x; t[1] += v.y; t[2] += v.z; t[3] += v.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = q * 4 + j;
            if (c == imax) t[j] += gres;
            if (c == imin) t[j] -= gres;
            t[j] *= dprelu_f(f[c], a);
        }
        reinterpret_cast<float4*>(gpre)[base + q * plane] = make_float4(t[0], t[1], t[2], t[3]);
    }
}

// gimg(q) = sum_t sum_c w[c][t] * gpre_c(q - t)
__global__ void __launch_bounds__(256)
stem_backward_kernel(const float* __restrict__ gpre, const float* __restrict__ w,
                     float* __restrict__ gimg, int H, int W) {
    __shared__ float4 sw[9 * STEM_C / 4];   // [tap][quad] -> 4 channels
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < 9 * STEM_C / 4; i += 256) {
        const int t = i / (STEM_C / 4), q = i % (STEM_C / 4);
        sw[i] = make_float4(w[(q * 4 + 0) * 9 + t], w[(q * 4 + 1) * 9 + t],
                            w[(q * 4 + 2) * 9 + t], w[(q * 4 + 3) * 9 + t]);
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float4* gp = reinterpret_cast<const float4*>(gpre) + (size_t)b * (STEM_C / 4) * plane;
    float acc = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y - dy, xx = x - dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const int t = (dy + 1) * 3 + (dx + 1);
#pragma unroll
            for (int q = 0; q < STEM_C / 4; ++q) {
                const float4 v = gp[q * plane + (size_t)yy * W + xx];
                const float4 ww = sw[t * (STEM_C / 4) + q];
                acc = fmaf(v.x, ww.x, acc); acc = fmaf(v.y, ww.y, acc);
                acc = fmaf(v.z, ww.z, acc); acc = fmaf(v.w, ww.w, acc);
            }
        }
    gimg[(size_t)b * plane + (size_t)y * W + x] = acc;
}

}  // namespace paif

using namespace paif;

extern "C" int paif_stem_forward(const float* img, long long sb, long long sy, long long sx,
                                 const float* w, const float* slope, float* feat, float* residue,
                                 int B, int H, int W, void* stream) {
    PAIF_REQUIRE(img && w && slope && feat && residue, "null pointer");
    PAIF_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B), block(32, 8);
    stem_forward_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(img, sb, 0, sy, sx, w, slope, feat, residue, nullptr, H, W);
    return check_launch("paif_stem_forward");
}

extern "C" int paif_stem_forward_bf16copy(const float* img, long long sb, long long sy, long long sx,
                                          const float* w, const float* slope, float* feat, float* residue,
                                          void* feat_bf16, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(img && w && slope && feat && residue && feat_bf16, "null pointer");
    PAIF_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B), block(32, 8);
    stem_forward_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(img, sb, 0, sy, sx, w, slope, feat, residue,
                                                                         static_cast<uint4*>(feat_bf16), H, W);
    return check_launch("paif_stem_forward_bf16copy");
}

extern "C" int paif_stem_forward_rgb(const float* img, long long sb, long long sc, long long sy, long long sx,
                                     const float* w, const float* slope, float* feat, float* residue,
                                     void* feat_bf16, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(img && w && slope && feat && residue, "null pointer");
    PAIF_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B), block(32, 8);
    stem_forward_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(img, sb, sc, sy, sx, w, slope, feat, residue,
                                                                        static_cast<uint4*>(feat_bf16), H, W);
    return check_launch("paif_stem_forward_rgb");
}

extern "C" int paif_stem_backward_pre(const float* feat, const float* slope,
                                      const float* g0, const float* g1, const float* g2, const float* g3,
                                      const float* gres_partial, int nparts, float* gpre,
                                      int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(feat && slope && gpre, "null pointer");
    PAIF_REQUIRE(!gres_partial || nparts > 0, "nparts must be positive");
    PAIF_REQUIRE(C == STEM_C, "C must be 32");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B), block(32, 8);
    stem_backward_pre_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(feat, slope, g0, g1, g2, g3,
                                                                     gres_partial, nparts, gpre, B, H, W);
    return check_launch("paif_stem_backward_pre");
}

extern "C" int paif_stem_backward(const float* gpre, const float* w, float* gimg,
                                  int C, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(gpre && w && gimg, "null pointer");
    PAIF_REQUIRE(C == STEM_C, "C must be 32");
    dim3 grid(cdiv(W, 32), cdiv(H, 8), B), block(32, 8);
    stem_backward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(gpre, w, gimg, H, W);
    return check_launch("paif_stem_backward");
}
