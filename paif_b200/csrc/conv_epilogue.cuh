// Fused convolution epilogue shared by the direct and the tcgen05 engines (see
// include/paif_b200.h, PaifConvDesc).  One thread owns one output pixel and all COUT channels.
#pragma once
#include "common.cuh"

namespace paif {

struct EpiParams {
    const float* ch_scale;
    const float* ch_shift;
    const float* pre_res[2];
    float* out_pre;
    const float* mask_src;
    const float* mask_slope;
    const float* slope;
    float post_scale;
    const float* post_res[3];
    float* out;
    float* out_act2;
    const float* slope2;
    float* chan_partials;
    int H, W;
};

inline EpiParams make_epi(const PaifConvDesc& d) {
    EpiParams e;
    e.ch_scale = d.ch_scale; e.ch_shift = d.ch_shift;
    // (map pointers: fp32 C4, or bf16 C8 when the launch's storage mode says so — the tcgen05 engine reinterprets)
    auto cf = [](const void* p) { return static_cast<const float*>(p); };
    e.pre_res[0] = cf(d.pre_res[0]); e.pre_res[1] = cf(d.pre_res[1]);
    e.out_pre = static_cast<float*>(d.out_pre); e.mask_src = cf(d.mask_src); e.mask_slope = d.mask_slope;
    e.slope = d.slope; e.post_scale = d.post_scale;
    e.post_res[0] = cf(d.post_res[0]); e.post_res[1] = cf(d.post_res[1]); e.post_res[2] = cf(d.post_res[2]);
    e.out = static_cast<float*>(d.out); e.out_act2 = static_cast<float*>(d.out_act2); e.slope2 = d.slope2;
    e.chan_partials = d.chan_partials;
    e.H = d.H; e.W = d.W;
    return e;
}

// v[COUT]: accumulators of pixel (b,y,x); on return v holds the stored values.
template <int COUT>
__device__ __forceinline__ void epilogue_pixel(const EpiParams& e, int b, int y, int x, float (&v)[COUT]) {
    constexpr int Q = COUT / 4;
    const size_t plane = (size_t)e.H * e.W;
    const size_t base = (size_t)b * Q * plane + (size_t)y * e.W + x;   // float4 units, quad 0
    const float a = e.slope ? *e.slope : 1.f;
    const float ma = e.mask_slope ? *e.mask_slope : 0.f;
    const float a2 = e.slope2 ? *e.slope2 : 1.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const size_t off = base + q * plane;
        float t[4] = {v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]};
        if (e.ch_scale) {
            const float4 s = reinterpret_cast<const float4*>(e.ch_scale)[q];
            t[0] *= s.x; t[1] *= s.y; t[2] *= s.z; t[3] *= s.w;
        }
        if (e.ch_shift) {
            const float4 s = reinterpret_cast<const float4*>(e.ch_shift)[q];
            t[0] += s.x; t[1] += s.y; t[2] += s.z; t[3] += s.w;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (e.pre_res[k]) {
                const float4 r = reinterpret_cast<const float4*>(e.pre_res[k])[off];
                t[0] += r.x; t[1] += r.y; t[2] += r.z; t[3] += r.w;
            }
        if (e.out_pre) reinterpret_cast<float4*>(e.out_pre)[off] = make_float4(t[0], t[1], t[2], t[3]);
        if (e.mask_src) {
            const float4 m = reinterpret_cast<const float4*>(e.mask_src)[off];
            t[0] *= dprelu_f(m.x, ma); t[1] *= dprelu_f(m.y, ma);
            t[2] *= dprelu_f(m.z, ma); t[3] *= dprelu_f(m.w, ma);
        } else if (e.slope) {
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = prelu_f(t[j], a);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] *= e.post_scale;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (e.post_res[k]) {
                const float4 r = reinterpret_cast<const float4*>(e.post_res[k])[off];
                t[0] += r.x; t[1] += r.y; t[2] += r.z; t[3] += r.w;
            }
        reinterpret_cast<float4*>(e.out)[off] = make_float4(t[0], t[1], t[2], t[3]);
        if (e.out_act2)
            reinterpret_cast<float4*>(e.out_act2)[off] =
                make_float4(prelu_f(t[0], a2), prelu_f(t[1], a2), prelu_f(t[2], a2), prelu_f(t[3], a2));
#pragma unroll
        for (int j = 0; j < 4; ++j) v[q * 4 + j] = t[j];
    }
}

}  // namespace paif
