// Whole-network entry point: Network_Fusion_Searched.forward(ir, vis) (core/model_fusion_auto.py:625-635) for the
// shipped `fusion_at` genotype (test_original.py:711-713) as ONE C-ABI call over a caller-owned workspace.
// It launches exactly the kernels — in exactly the order — that the Python orchestration of paif_b200/fusion.py
// launches for that genotype on the tcgen05 engine (fused decomposition, dense DilConv, stem_out on the engine), so
// the two are bit-identical; what disappears is ~27 ctypes calls, a tensor allocation per buffer and the Python
// between them (batch-1 latency).  No allocation, no synchronisation: CUDA-graph capturable like every other entry.
#include "common.cuh"

using namespace paif;

namespace {

struct Arena {
    unsigned char* base;
    size_t off, cap;
    bool dry;
    void* take(size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        void* p = dry ? nullptr : base + off;
        off += bytes;
        return p;
    }
};

struct Ctx {
    const PaifFusionWeights* w;
    int B, H, W, bf16;
    void* stream;
    size_t map_bytes() const { return (size_t)B * H * W * 32 * (bf16 ? 2 : 4); }
    size_t map32_bytes() const { return (size_t)B * H * W * 32 * 4; }
    size_t plane_bytes() const { return (size_t)B * H * W * 4; }
};

int conv(const Ctx& c, const PaifFusionConv& cw, int nsrc, int k, int dil, const void* s0, const void* s1, const void* s2,
         void* out, const float* slope, float post_scale, const void* r0, const void* r1, const void* r2,
         const float* ch_scale = nullptr, const float* ch_shift = nullptr, void* act2 = nullptr, const float* slope2 = nullptr,
         float* partials = nullptr) {
    PaifConvDesc d = {};
    d.B = c.B; d.H = c.H; d.W = c.W;
    d.nsrc = nsrc; d.cin_per_src = 32; d.cout = 32; d.kh = d.kw = k; d.dil = dil;
    d.engine = PAIF_ENGINE_TCGEN05;
    d.src[0] = s0; d.src[1] = s1; d.src[2] = s2;
    d.weight = cw.direct;
    d.weight_mma = c.bf16 ? cw.mma_bf16 : cw.mma_tf32;
    d.ch_scale = ch_scale; d.ch_shift = ch_shift;
    d.slope = slope; d.post_scale = post_scale;
    d.post_res[0] = r0; d.post_res[1] = r1; d.post_res[2] = r2;
    d.out = out; d.out_act2 = act2; d.slope2 = slope2; d.chan_partials = partials;
    d.storage = c.bf16 ? PAIF_STORAGE_BF16 : PAIF_STORAGE_F32;
    return paif_conv_forward(&d, c.stream);
}

// ResidualDenseBlock (operations_m.py:435-449): out = PReLU(c3([x, x1, x2])) * 0.333333 + x (+ r1 + r2)
int rdb(const Ctx& c, const PaifFusionRDB& p, const void* x, void* x1, void* x2, void* out, const void* r1, const void* r2,
        void* relu_out, const float* zero_slope) {
    if (int r = conv(c, p.conv[0], 1, 3, 1, x, nullptr, nullptr, x1, p.slope, 1.f, nullptr, nullptr, nullptr)) return r;
    if (int r = conv(c, p.conv[1], 2, 3, 1, x, x1, nullptr, x2, p.slope, 1.f, nullptr, nullptr, nullptr)) return r;
    return conv(c, p.conv[2], 3, 3, 1, x, x1, x2, out, p.slope, 0.333333f, x, r1, r2, nullptr, nullptr,
                relu_out, relu_out ? zero_slope : nullptr);
}

int run(const PaifFusionWeights* w, const float* ir, long long ir_sb, long long ir_sy, long long ir_sx,
        const float* vis, long long vis_sb, long long vis_sy, long long vis_sx, float* out,
        void* workspace, size_t workspace_bytes, int storage, int B, int H, int W, void* stream, size_t* need) {
    Ctx c{w, B, H, W, storage == PAIF_STORAGE_BF16, stream};
    Arena a{static_cast<unsigned char*>(workspace), 0, workspace_bytes, need != nullptr};
    const size_t mb = c.map_bytes(), m32 = c.map32_bytes(), pb = c.plane_bytes();
    // buffers (a handful of map-sized slots reused along the chain; the stem features live until their branch ends)
    float* feat[2] = {static_cast<float*>(a.take(m32)), static_cast<float*>(a.take(m32))};
    void* feat16[2] = {c.bf16 ? a.take(mb) : nullptr, c.bf16 ? a.take(mb) : nullptr};
    float* guide[2] = {static_cast<float*>(a.take(pb)), static_cast<float*>(a.take(pb))};
    float* stats = static_cast<float*>(a.take(3 * pb));
    void* s[6];
    for (int i = 0; i < 6; ++i) s[i] = a.take(mb);
    void* branch[2] = {a.take(mb), a.take(mb)};
    const int tiles = paif_conv_num_tiles(H, W, PAIF_ENGINE_TCGEN05);
    float* partials = static_cast<float*>(a.take((size_t)B * tiles * 32 * sizeof(float)));
    float* eca_e = static_cast<float*>(a.take((size_t)B * 32 * sizeof(float)));
    float* zero = static_cast<float*>(a.take(256));
    if (need) { *need = a.off; return 0; }
    if (a.off > workspace_bytes) { set_error("paif_fusion_forward: workspace too small (%zu > %zu bytes)", a.off, workspace_bytes); return PAIF_EINVAL; }
    cudaError_t ce = cudaMemsetAsync(zero, 0, 256, (cudaStream_t)stream);
    if (ce != cudaSuccess) { set_error("paif_fusion_forward: memset: %s", cudaGetErrorString(ce)); return (int)ce; }

    // stems + guide (core/model_fusion_auto.py:628-629, 517-521)
    const float* img[2] = {ir, vis};
    const long long sb[2] = {ir_sb, vis_sb}, sy[2] = {ir_sy, vis_sy}, sx[2] = {ir_sx, vis_sx};
    for (int i = 0; i < 2; ++i) {
        int r = c.bf16 ? paif_stem_forward_bf16copy(img[i], sb[i], sy[i], sx[i], w->stem_w[i], w->stem_a[i], feat[i], guide[i], feat16[i], B, H, W, stream)
                       : paif_stem_forward(img[i], sb[i], sy[i], sx[i], w->stem_w[i], w->stem_a[i], feat[i], guide[i], B, H, W, stream);
        if (r) return r;
    }
    // decomposition branches (:509-516): IR = RDB -> DilConv, VIS = RDB -> RDB, each + chain input + stem feature
    for (int i = 0; i < 2; ++i) {
        const void* fres = c.bf16 ? feat16[i] : static_cast<const void*>(feat[i]);
        if (int r = paif_gf_guide_stats(guide[i], stats, B, H, W, stream)) return r;
        void* x = s[0];
        if (int r = paif_gf_mix_forward(feat[i], guide[i], stats, w->gfmix_w[i], w->c1x1_b[i], x, c.bf16, 32, B, H, W, stream)) return r;
        if (i == 0) {
            // chain: lf + DilConv_3_2(Denseblocks_3_1(lf)); DilConv as one dense 3x3 (dil 2) convolution over relu(RDB out)
            if (int r = rdb(c, w->rdb[0], x, s[1], s[2], s[3], nullptr, nullptr, s[4], zero)) return r;
            if (int r = conv(c, w->dil_dense, 1, 3, 2, s[4], nullptr, nullptr, branch[0], nullptr, 1.f, s[3], x, fres,
                             w->dil_scale, w->dil_shift)) return r;
        } else {
            if (int r = rdb(c, w->rdb[1], x, s[1], s[2], s[3], nullptr, nullptr, nullptr, zero)) return r;
            if (int r = rdb(c, w->rdb[2], s[3], s[1], s[2], branch[1], x, fres, nullptr, zero)) return r;
        }
    }
    // spatial attention blend (:631-632, 1352-1368)
    void* agg = s[0];
    {
        int r = c.bf16 ? paif_spa_fused_forward_bf16(w->spa_w, w->spa_k, branch[0], branch[1], agg, 32, B, H, W, stream)
                       : paif_spa_fused_forward(w->spa_w, w->spa_k, static_cast<const float*>(branch[0]), static_cast<const float*>(branch[1]),
                                                static_cast<float*>(agg), nullptr, 32, B, H, W, stream);
        if (r) return r;
    }
    // final chain (:633): agg + Residualblocks_7_1(ECAattention_3(agg))
    void *x0 = s[1], *px0 = s[2], *o = s[3], *eca_out = s[4], *t1 = s[5], *f2 = branch[0];
    if (int r = conv(c, w->eca_conv1, 1, 3, 1, agg, nullptr, nullptr, x0, nullptr, 1.f, nullptr, nullptr, nullptr, nullptr, nullptr, px0, w->eca_a)) return r;
    if (int r = conv(c, w->eca_conv2, 1, 3, 1, px0, nullptr, nullptr, o, nullptr, 1.f, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, partials)) return r;
    if (int r = paif_eca_scale(partials, tiles, w->eca_w1d, 3, eca_e, 32, B, H, W, stream)) return r;
    {
        int r = c.bf16 ? paif_eca_apply_bf16(o, x0, eca_e, w->eca_a, nullptr, eca_out, 32, B, H, W, stream)
                       : paif_eca_apply(static_cast<const float*>(o), static_cast<const float*>(x0), eca_e, w->eca_a, nullptr,
                                        static_cast<float*>(eca_out), 32, B, H, W, stream);
        if (r) return r;
    }
    if (int r = conv(c, w->res_conv7, 1, 7, 1, eca_out, nullptr, nullptr, t1, nullptr, 1.f, nullptr, nullptr, nullptr)) return r;
    if (int r = conv(c, w->res_merged, 1, 3, 2, t1, nullptr, nullptr, f2, w->res_a, 1.f, eca_out, agg, nullptr, w->res_scale, w->res_shift)) return r;
    // stem_out + tanh (:615-620, 634)
    return paif_out_forward_tc(f2, c.bf16 ? w->out_mma_bf16 : w->out_mma_tf32, w->out_wm, w->out_a, out, nullptr,
                               c.bf16 ? PAIF_STORAGE_BF16 : PAIF_STORAGE_F32, 32, B, H, W, stream);
}

}  // namespace

extern "C" long long paif_fusion_workspace_bytes(int B, int H, int W, int storage) {
    if (B <= 0 || H <= 9 || W <= 9 || (storage != PAIF_STORAGE_F32 && storage != PAIF_STORAGE_BF16)) return -1;
    size_t need = 0;
    run(nullptr, nullptr, 0, 0, 0, nullptr, 0, 0, 0, nullptr, nullptr, 0, storage, B, H, W, nullptr, &need);
    return (long long)need;
}

extern "C" int paif_fusion_forward(const PaifFusionWeights* weights,
                                   const float* ir, long long ir_stride_b, long long ir_stride_y, long long ir_stride_x,
                                   const float* vis, long long vis_stride_b, long long vis_stride_y, long long vis_stride_x,
                                   float* out, void* workspace, long long workspace_bytes, int storage,
                                   int B, int H, int W, void* stream) {
    PAIF_REQUIRE(weights && ir && vis && out && workspace, "null pointer");
    PAIF_REQUIRE(B > 0 && B <= 65535 && H > 9 && W > 9, "bad shape (the guided filter needs H, W > 9)");
    PAIF_REQUIRE(storage == PAIF_STORAGE_F32 || storage == PAIF_STORAGE_BF16, "storage must be F32 or BF16");
    PAIF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    if (!paif_gf_mix_supported(32, H, W)) { set_error("paif_fusion_forward: needs W %% 4 == 0"); return PAIF_ENOTSUP; }
    return run(weights, ir, ir_stride_b, ir_stride_y, ir_stride_x, vis, vis_stride_b, vis_stride_y, vis_stride_x, out,
               workspace, (size_t)workspace_bytes, storage, B, H, W, stream, nullptr);
}
