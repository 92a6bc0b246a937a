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

// ---------------------------------------------------------------------------------------------
// Forward with saved activations + backward-to-input (attack/attack.py:444-501: the PGD inner step needs d loss / d ir and
// d loss / d vis) as two calls over ONE caller-owned workspace: the forward leaves the activations the backward needs at
// fixed offsets of it, the backward uses the rest as scratch.  fp32 storage (the PGD loop's mode).  Same kernels, same
// order as paif_b200/fusion.py::_FusionFn (per-operator path), so gradients are bit-identical to it.
// ---------------------------------------------------------------------------------------------
struct TrainBufs {
    float *feat[2], *guide[2], *gstats[2];
    float *gf_ma[2];                        // mean2(A') of the fused decomposition (direct guide term of its adjoint)
    float *rx1[3], *rx2[3], *rpre3[3];      // RDB records: IR chain RDB, VIS chain RDB 1 / 2
    float *s1;                              // IR chain: RDB output (DilConv's input: its ReLU' mask source)
    float *a_f, *v_f, *scale;               // branch outputs, attention plane
    float *eca_x0, *eca_o, *eca_e, *res_pre, *pre_out, *out_copy;
    float *t[8];                            // scratch maps (forward temporaries / backward gradients)
    float *partials, *bpartials, *gm, *gres, *work, *zero, *plane[2];
    int tiles, btiles, nparts;
};

size_t train_layout(int B, int H, int W, unsigned char* base, TrainBufs* tb) {
    Arena a{base, 0, 0, base == nullptr};
    const size_t m = (size_t)B * H * W * 32 * 4, pl = (size_t)B * H * W * 4;
    auto map = [&]() { return static_cast<float*>(a.take(m)); };
    auto plane = [&](int n = 1) { return static_cast<float*>(a.take(n * pl)); };
    for (int i = 0; i < 2; ++i) { tb->feat[i] = map(); tb->guide[i] = plane(); tb->gstats[i] = plane(3); tb->gf_ma[i] = map(); }
    for (int i = 0; i < 3; ++i) { tb->rx1[i] = map(); tb->rx2[i] = map(); tb->rpre3[i] = map(); }
    tb->s1 = map(); tb->a_f = map(); tb->v_f = map(); tb->scale = plane();
    tb->eca_x0 = map(); tb->eca_o = map(); tb->res_pre = map(); tb->pre_out = plane(); tb->out_copy = plane();
    tb->eca_e = static_cast<float*>(a.take((size_t)B * 32 * sizeof(float)));
    for (int i = 0; i < 8; ++i) tb->t[i] = map();
    tb->tiles = paif_conv_num_tiles(H, W, PAIF_ENGINE_TCGEN05);
    tb->btiles = paif_eca_bwd_tiles(H, W);
    tb->nparts = paif_gf_guide_parts(32);
    tb->partials = static_cast<float*>(a.take((size_t)B * tb->tiles * 32 * sizeof(float)));
    tb->bpartials = static_cast<float*>(a.take((size_t)B * tb->btiles * 32 * sizeof(float)));
    tb->gm = static_cast<float*>(a.take((size_t)B * 32 * sizeof(float)));
    tb->gres = plane(tb->nparts);
    tb->work = static_cast<float*>(a.take((size_t)paif_gf_backward_work_floats(32, B, H, W) * sizeof(float)));
    tb->plane[0] = plane(); tb->plane[1] = plane();
    tb->zero = static_cast<float*>(a.take(256));
    return a.off;
}

// one convolution on the tcgen05 engine, fp32 storage, with the full epilogue vocabulary
struct CArgs {
    const void* src[3] = {nullptr, nullptr, nullptr};
    const float* slope = nullptr;
    float post_scale = 1.f;
    const void* post[3] = {nullptr, nullptr, nullptr};
    const void* pre[2] = {nullptr, nullptr};
    const void* mask_src = nullptr;
    const float* mask_slope = nullptr;
    const float *ch_scale = nullptr, *ch_shift = nullptr;
    void *out_pre = nullptr, *act2 = nullptr;
    const float* slope2 = nullptr;
    float* partials = nullptr;
};
int convx(const Ctx& c, const PaifFusionConv& cw, int nsrc, int k, int dil, void* out, const CArgs& x) {
    PaifConvDesc d = {};
    d.B = c.B; d.H = c.H; d.W = c.W;
    d.nsrc = nsrc; d.cin_per_src = 32; d.cout = 32; d.kh = d.kw = k; d.dil = dil;
    d.engine = PAIF_ENGINE_TCGEN05;
    for (int i = 0; i < 3; ++i) { d.src[i] = x.src[i]; d.post_res[i] = x.post[i]; }
    d.pre_res[0] = x.pre[0]; d.pre_res[1] = x.pre[1];
    d.weight = cw.direct; d.weight_mma = cw.mma_tf32;
    d.ch_scale = x.ch_scale; d.ch_shift = x.ch_shift;
    d.slope = x.slope; d.post_scale = x.post_scale;
    d.mask_src = x.mask_src; d.mask_slope = x.mask_slope;
    d.out = out; d.out_pre = x.out_pre; d.out_act2 = x.act2; d.slope2 = x.slope2; d.chan_partials = x.partials;
    d.storage = PAIF_STORAGE_F32;
    return paif_conv_forward(&d, c.stream);
}
#define TRY(expr) do { if (int r_ = (expr)) return r_; } while (0)

// ResidualDenseBlock forward with its records (x1, x2, pre3)
int rdb_save(const Ctx& c, const PaifFusionRDB& p, const float* x, float* x1, float* x2, float* pre3, float* out,
             const void* r1, const void* r2, float* relu_out, const float* zero) {
    CArgs a1; a1.src[0] = x; a1.slope = p.slope;
    TRY(convx(c, p.conv[0], 1, 3, 1, x1, a1));
    CArgs a2; a2.src[0] = x; a2.src[1] = x1; a2.slope = p.slope;
    TRY(convx(c, p.conv[1], 2, 3, 1, x2, a2));
    CArgs a3; a3.src[0] = x; a3.src[1] = x1; a3.src[2] = x2; a3.slope = p.slope; a3.post_scale = 0.333333f;
    a3.post[0] = x; a3.post[1] = r1; a3.post[2] = r2; a3.out_pre = pre3; a3.act2 = relu_out; a3.slope2 = relu_out ? zero : nullptr;
    return convx(c, p.conv[2], 3, 3, 1, out, a3);
}

// ResidualDenseBlock backward (fusion.py::ResidualDenseBlock.bwd): g -> gradient w.r.t. the block input (+ e0 + e1)
int rdb_bwd(const Ctx& c, const PaifFusionRDBGrad& wd, const float* slope, const float* x1, const float* x2, const float* pre3,
            const float* g, const float* e0, float* out, float* const* t /* 6 scratch maps */) {
    float *g3 = t[0], *gxa = t[1], *gx1a = t[2], *g2 = t[3], *gxb = t[4], *g1 = t[5];
    TRY(paif_mask_scale(g, pre3, slope, 0.333333f, g3, 32, c.B, c.H, c.W, c.stream));
    { CArgs a; a.src[0] = g3; TRY(convx(c, wd.c3[0], 1, 3, 1, gxa, a)); }
    { CArgs a; a.src[0] = g3; TRY(convx(c, wd.c3[1], 1, 3, 1, gx1a, a)); }
    { CArgs a; a.src[0] = g3; a.mask_src = x2; a.mask_slope = slope; TRY(convx(c, wd.c3[2], 1, 3, 1, g2, a)); }
    { CArgs a; a.src[0] = g2; a.post[0] = gxa; TRY(convx(c, wd.c2[0], 1, 3, 1, gxb, a)); }
    { CArgs a; a.src[0] = g2; a.pre[0] = gx1a; a.mask_src = x1; a.mask_slope = slope; TRY(convx(c, wd.c2[1], 1, 3, 1, g1, a)); }
    CArgs a; a.src[0] = g1; a.post[0] = gxb; a.post[1] = g; a.post[2] = e0;
    return convx(c, wd.c1, 1, 3, 1, out, a);
}

int train_forward(const PaifFusionWeights* w, const float* ir, long long ir_sb, long long ir_sy, long long ir_sx,
                  const float* vis, long long vis_sb, long long vis_sy, long long vis_sx, float* out,
                  void* workspace, int B, int H, int W, void* stream) {
    Ctx c{w, B, H, W, 0, stream};
    TrainBufs tb;
    train_layout(B, H, W, static_cast<unsigned char*>(workspace), &tb);
    cudaError_t ce = cudaMemsetAsync(tb.zero, 0, 256, (cudaStream_t)stream);
    if (ce != cudaSuccess) { set_error("paif_fusion_forward_save: memset: %s", cudaGetErrorString(ce)); return (int)ce; }
    const float* img[2] = {ir, vis};
    const long long sb[2] = {ir_sb, vis_sb}, sy[2] = {ir_sy, vis_sy}, sx[2] = {ir_sx, vis_sx};
    for (int i = 0; i < 2; ++i)
        TRY(paif_stem_forward(img[i], sb[i], sy[i], sx[i], w->stem_w[i], w->stem_a[i], tb.feat[i], tb.guide[i], B, H, W, stream));
    for (int i = 0; i < 2; ++i) {
        TRY(paif_gf_guide_stats(tb.guide[i], tb.gstats[i], B, H, W, stream));
        float* x = tb.t[0];
        TRY(paif_gf_mix_forward_save(tb.feat[i], tb.guide[i], tb.gstats[i], w->gfmix_w[i], w->c1x1_b[i], x, 0, tb.gf_ma[i], 32, B, H, W, stream));
        if (i == 0) {
            TRY(rdb_save(c, w->rdb[0], x, tb.rx1[0], tb.rx2[0], tb.rpre3[0], tb.s1, nullptr, nullptr, tb.t[1], tb.zero));
            CArgs a; a.src[0] = tb.t[1]; a.post[0] = tb.s1; a.post[1] = x; a.post[2] = tb.feat[0];
            a.ch_scale = w->dil_scale; a.ch_shift = w->dil_shift;
            TRY(convx(c, w->dil_dense, 1, 3, 2, tb.a_f, a));
        } else {
            TRY(rdb_save(c, w->rdb[1], x, tb.rx1[1], tb.rx2[1], tb.rpre3[1], tb.t[1], nullptr, nullptr, nullptr, tb.zero));
            TRY(rdb_save(c, w->rdb[2], tb.t[1], tb.rx1[2], tb.rx2[2], tb.rpre3[2], tb.v_f, x, tb.feat[1], nullptr, tb.zero));
        }
    }
    float* agg = tb.t[0];
    TRY(paif_spa_fused_forward(w->spa_w, w->spa_k, tb.a_f, tb.v_f, agg, tb.scale, 32, B, H, W, stream));
    float *px0 = tb.t[1], *eca_out = tb.t[2], *t1 = tb.t[3], *f2 = tb.t[4];
    { CArgs a; a.src[0] = agg; a.act2 = px0; a.slope2 = w->eca_a; TRY(convx(c, w->eca_conv1, 1, 3, 1, tb.eca_x0, a)); }
    { CArgs a; a.src[0] = px0; a.partials = tb.partials; TRY(convx(c, w->eca_conv2, 1, 3, 1, tb.eca_o, a)); }
    TRY(paif_eca_scale(tb.partials, tb.tiles, w->eca_w1d, 3, tb.eca_e, 32, B, H, W, stream));
    TRY(paif_eca_apply(tb.eca_o, tb.eca_x0, tb.eca_e, w->eca_a, nullptr, eca_out, 32, B, H, W, stream));
    { CArgs a; a.src[0] = eca_out; TRY(convx(c, w->res_conv7, 1, 7, 1, t1, a)); }
    { CArgs a; a.src[0] = t1; a.slope = w->res_a; a.post[0] = eca_out; a.post[1] = agg; a.ch_scale = w->res_scale; a.ch_shift = w->res_shift;
      a.out_pre = tb.res_pre; TRY(convx(c, w->res_merged, 1, 3, 2, f2, a)); }
    TRY(paif_out_forward_tc(f2, w->out_mma_tf32, w->out_wm, w->out_a, out, tb.pre_out, PAIF_STORAGE_F32, 32, B, H, W, stream));
    // the backward needs tanh's output: keep a copy inside the workspace (the caller may overwrite `out`)
    ce = cudaMemcpyAsync(tb.out_copy, out, (size_t)B * H * W * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (ce != cudaSuccess) { set_error("paif_fusion_forward_save: copy: %s", cudaGetErrorString(ce)); return (int)ce; }
    return 0;
}

int train_backward(const PaifFusionWeights* w, const PaifFusionGradWeights* wd, const float* gout, float* g_ir, float* g_vis,
                   void* workspace, int B, int H, int W, void* stream) {
    Ctx c{w, B, H, W, 0, stream};
    TrainBufs tb;
    train_layout(B, H, W, static_cast<unsigned char*>(workspace), &tb);
    float** t = tb.t;
    cudaError_t ce = cudaMemsetAsync(tb.zero, 0, 256, (cudaStream_t)stream);
    if (ce != cudaSuccess) { set_error("paif_fusion_backward_input: memset: %s", cudaGetErrorString(ce)); return (int)ce; }
    // stem_out + tanh, with the trailing ResidualModule's PReLU' mask fused (gmasked)
    float *gf2 = t[0], *gmasked = t[1];
    TRY(paif_out_backward(gout, tb.out_copy, tb.pre_out, w->out_wm, w->out_a, gf2, tb.res_pre, w->res_a, gmasked, 32, B, H, W, stream));
    // ResidualModule: merged 3x3(dil 2) dgrad (BatchNorm scale folded), 7x7 dgrad + the block's own skip
    float *gt1 = t[2], *gs = t[3];
    { CArgs a; a.src[0] = gmasked; TRY(convx(c, wd->res_merged_d, 1, 3, 2, gt1, a)); }
    { CArgs a; a.src[0] = gt1; a.post[0] = gf2; TRY(convx(c, wd->res_conv7_d, 1, 7, 1, gs, a)); }
    // ECABasicBlock
    float *gw = t[1], *go = t[2], *gx0 = t[4], *g_agg = t[5];
    TRY(paif_eca_bwd_pass1(gs, tb.eca_o, tb.eca_x0, tb.eca_e, w->eca_a, gw, tb.bpartials, 32, B, H, W, stream));
    TRY(paif_eca_bwd_scale(tb.bpartials, tb.btiles, tb.eca_e, w->eca_w1d, 3, tb.gm, 32, B, H, W, stream));
    TRY(paif_eca_bwd_pass2(gw, tb.eca_e, tb.gm, go, 32, B, H, W, stream));
    { CArgs a; a.src[0] = go; a.mask_src = tb.eca_x0; a.mask_slope = w->eca_a; a.post[0] = gw; TRY(convx(c, wd->eca_conv2_d, 1, 3, 1, gx0, a)); }
    { CArgs a; a.src[0] = gx0; a.post[0] = gf2; TRY(convx(c, wd->eca_conv1_d, 1, 3, 1, g_agg, a)); }
    // blend + spatial attention + ChannelPool
    float *gpre = tb.plane[0], *g_a = t[6], *g_v = t[7];
    TRY(paif_spa_blend_backward_pre(g_agg, tb.a_f, tb.v_f, tb.scale, gpre, 32, B, H, W, stream));
    TRY(paif_spa_blend_backward(g_agg, tb.a_f, tb.v_f, tb.scale, gpre, w->spa_w, w->spa_k, g_a, g_v, 32, B, H, W, stream));
    // branches.  Scratch: t[0..5] and the records of the final chain, which are dead by now (eca_o, res_pre); t[6] / t[7]
    // hold the two branch gradients until their branch is done.
    float* gimg[2] = {g_ir, g_vis};
    float* rs[6] = {t[0], t[1], t[2], t[3], t[4], tb.eca_o};       // the six temporaries of a dense-block backward
    float* gx = tb.res_pre;                                        // gradient w.r.t. the fused decomposition's output
    for (int i = 0; i < 2; ++i) {
        const float* gb = i == 0 ? g_a : g_v;
        if (i == 0) {
            // DilConv: one dense dgrad convolution, ReLU' of its input as the mask, its "+x" skip as a residual
            float* gsd = t[5];
            { CArgs a; a.src[0] = gb; a.mask_src = tb.s1; a.mask_slope = tb.zero; a.post[0] = gb; TRY(convx(c, wd->dil_dense_d, 1, 3, 2, gsd, a)); }
            TRY(rdb_bwd(c, wd->rdb[0], w->rdb[0].slope, tb.rx1[0], tb.rx2[0], tb.rpre3[0], gsd, gb, gx, rs));
        } else {
            float* gs2 = t[5];
            TRY(rdb_bwd(c, wd->rdb[2], w->rdb[2].slope, tb.rx1[2], tb.rx2[2], tb.rpre3[2], gb, nullptr, gs2, rs));
            TRY(rdb_bwd(c, wd->rdb[1], w->rdb[1].slope, tb.rx1[1], tb.rx2[1], tb.rpre3[1], gs2, gb, gx, rs));
        }
        // the folded 1x1 of the fused decomposition, differentiated as its three K groups; then the guided-filter adjoint
        float *glf1 = t[0], *glf2 = t[1], *gz = t[2], *gfeat = t[3];
        { CArgs a; a.src[0] = gx; TRY(convx(c, wd->c1x1_d[i][0], 1, 1, 1, glf1, a)); }
        { CArgs a; a.src[0] = gx; TRY(convx(c, wd->c1x1_d[i][1], 1, 1, 1, glf2, a)); }
        { CArgs a; a.src[0] = gx; TRY(convx(c, wd->c1x1_d[i][2], 1, 1, 1, gz, a)); }
        TRY(paif_gf_decomp_backward_saved(tb.feat[i], tb.guide[i], tb.gstats[i], glf1, glf2, gx, tb.gf_ma[i], gfeat, tb.gres, tb.work, 32, B, H, W, stream));
        float* gstem = t[0];
        TRY(paif_stem_backward_pre(tb.feat[i], w->stem_a[i], gb, gz, gfeat, nullptr, tb.gres, tb.nparts, gstem, 32, B, H, W, stream));
        TRY(paif_stem_backward(gstem, w->stem_w[i], gimg[i], 32, B, H, W, stream));
    }
    return 0;
}

}  // namespace

extern "C" long long paif_fusion_train_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 9 || W <= 9) return -1;
    TrainBufs tb;
    return (long long)train_layout(B, H, W, nullptr, &tb);
}

extern "C" int paif_fusion_forward_save(const PaifFusionWeights* weights,
                                        const float* ir, long long ir_stride_b, long long ir_stride_y, long long ir_stride_x,
                                        const float* vis, long long vis_stride_b, long long vis_stride_y, long long vis_stride_x,
                                        float* out, void* workspace, long long workspace_bytes, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(weights && ir && vis && out && workspace, "null pointer");
    PAIF_REQUIRE(B > 0 && B <= 65535 && H > 9 && W > 9, "bad shape (the guided filter needs H, W > 9)");
    PAIF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    if (!paif_gf_mix_supported(32, H, W)) { set_error("paif_fusion_forward_save: needs W %% 4 == 0"); return PAIF_ENOTSUP; }
    PAIF_REQUIRE(workspace_bytes >= paif_fusion_train_workspace_bytes(B, H, W), "workspace too small");
    return train_forward(weights, ir, ir_stride_b, ir_stride_y, ir_stride_x, vis, vis_stride_b, vis_stride_y, vis_stride_x, out,
                         workspace, B, H, W, stream);
}

extern "C" int paif_fusion_backward_input(const PaifFusionWeights* weights, const PaifFusionGradWeights* grad_weights,
                                          const float* gout, float* g_ir, float* g_vis,
                                          void* workspace, long long workspace_bytes, int B, int H, int W, void* stream) {
    PAIF_REQUIRE(weights && grad_weights && gout && g_ir && g_vis && workspace, "null pointer");
    PAIF_REQUIRE(B > 0 && B <= 65535 && H > 9 && W > 9, "bad shape");
    PAIF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    PAIF_REQUIRE(workspace_bytes >= paif_fusion_train_workspace_bytes(B, H, W), "workspace too small");
    return train_backward(weights, grad_weights, gout, g_ir, g_vis, workspace, B, H, W, stream);
}

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
