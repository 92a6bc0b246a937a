/*
 * paif_b200 — C ABI of the B200-native PAIF fusion hot path.
 *
 * The reference (LiuZhu-CV/PAIF) has no FFI of its own: its boundary for this path is
 * the Python nn.Module interface of Network_Fusion_Searched
 * (core/model_fusion_auto.py:599-640).  This header is the C-ABI layer underneath the
 * drop-in module (paif_b200/fusion.py); every entry point names the reference
 * operator it replaces.  All functions
 *   - take plain device pointers, ints and a cudaStream_t passed as void*;
 *   - enqueue on the caller's stream, never allocate, never synchronise
 *     (so a whole forward / backward is CUDA-graph capturable);
 *   - return 0 on success, a negative PAIF_E* validation code, or a positive
 *     cudaError_t from the launch; paif_last_error_string() describes the last failure
 *     of the calling thread.
 * The caller (PyTorch) owns every buffer.
 *
 * ACTIVATION LAYOUT ("C4 map"): a C-channel fp32 feature map of a batch of B images
 * is stored as [B][C/4][H][W][4] — 16-byte channel quads, planar per quad.  A warp
 * that walks x reads/writes 512 contiguous bytes per quad plane, and one quad plane
 * row is exactly the K-major, no-swizzle UMMA core-matrix operand (8 pixels x 16 B)
 * that the tcgen05 implicit-GEMM convolution streams through shared memory.
 * 1-channel planes are [B][H][W]; the pooled attention input is [B][H][W][4].
 *
 * Scalars that are learnable parameters (PReLU slopes) are passed as DEVICE pointers so
 * that no host read-back is ever needed.
 */
#ifndef PAIF_B200_H
#define PAIF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PAIF_ABI_VERSION 5   /* 5: paif_gf_mix_forward_save / paif_gf_decomp_backward_saved (direct guide term of the adjoint from the forward's mean2(A')); 4: paif_fusion_forward_save / paif_fusion_backward_input, paif_conv_set_persistent; 3: paif_gf_mix_forward, glue / PGD / loss-head kernels, paif_stem_forward_rgb, paif_widen_bf16_map */

#define PAIF_EINVAL   (-1)   /* bad argument (null pointer, unsupported size) */
#define PAIF_ENOTSUP  (-2)   /* configuration not supported by this build     */

/* conv engines */
/* Map storage of a convolution launch (PaifConvDesc.storage).  fp32 maps use the C4 layout [B][C/4][H][W][4];
 * bf16 maps the C8 layout [B][C/8][H][W][8] (16-byte pixel vectors either way).  The bf16 modes exist on the
 * tcgen05 engine only and are forward-only in the module (north_star's 1e-2 tier; fp32 accumulation in TMEM,
 * per-channel affine / PReLU / residual adds in fp32 registers, one rounding at the store). */
#define PAIF_STORAGE_F32       0   /* fp32 sources (TF32 operands), fp32 residuals and outputs            */
#define PAIF_STORAGE_BF16      1   /* bf16 sources (bf16 operands, kind::f16), bf16 residuals and outputs */
#define PAIF_STORAGE_F32_BF16  2   /* fp32 sources (TF32 operands), bf16 residuals and outputs (1x1 only) */

#define PAIF_ENGINE_AUTO    0
#define PAIF_ENGINE_DIRECT  1   /* fp32 FFMA direct convolution (exact-fp32 path)          */
#define PAIF_ENGINE_TCGEN05 2   /* tcgen05 TF32 implicit GEMM, fp32 accumulate in TMEM      */

int         paif_abi_version(void);
const char* paif_last_error_string(void);

/* ------------------------------------------------------------------------------------
 * stem_1 / stem_2: Conv2d(1,32,3,pad 1,bias=False) + PReLU(1)  — core/model_fusion_auto.py:607-614,
 * fused with Cell_Decom.get_residue (max_c - min_c) — :517-521.
 * img: 1-channel image with arbitrary element strides (the wrappers pass a channel-last
 * strided view for vis).  w: [32][9].  feat: C4 map (32 ch).  residue: [B][H][W]. */
int paif_stem_forward(const float* img, long long stride_b, long long stride_y, long long stride_x,
                      const float* w, const float* slope,
                      float* feat, float* residue, int B, int H, int W, void* stream);

/* paif_stem_forward reading the VISIBLE image as RGB and forming Y = .299 R + .587 G + .114 B (RGB2YCrCb,
 * core/model_fusion_auto.py:69-92) on the fly: img points at the R plane, stride_c is the channel stride. */
int paif_stem_forward_rgb(const float* img, long long stride_b, long long stride_c, long long stride_y, long long stride_x,
                          const float* w, const float* slope,
                          float* feat, float* residue, void* feat_bf16, int B, int H, int W, void* stream);

/* GuidedFilter(4, eps)(residue, feat) for eps in {1e-3, 1e-4} — core/model_fusion_auto.py:522-535
 * + third-party guided_filter_pytorch (box filter radius 4, clipped borders).
 * Pass 0 computes what all 32 channels share: stats = [3][B][H][W] = mean_guide, 1/(var_guide + 1e-3),
 * 1/(var_guide + 1e-4) (caller-owned scratch).  Pass 1 writes the two low-frequency maps;
 * HF = feat - LF is folded into the 1x1 conv weights. */
int paif_gf_guide_stats(const float* residue, float* stats, int B, int H, int W, void* stream);
int paif_gf_decomp_forward(const float* feat, const float* residue, const float* stats,
                           float* lf1, float* lf2, int C, int B, int H, int W, void* stream);

/* Fused decomposition + folded 1x1: Cell_Decom.decomposition followed by conv1x1_lf / conv1x1_hf
 * (core/model_fusion_auto.py:509-535) in ONE kernel that never writes the LF maps:
 *     out = Wa LF_1e-3 + Wb LF_1e-4 + Wc feat + bias,
 * Wa = W[:,0:32]-W[:,64:96], Wb = W[:,32:64]-W[:,96:128], Wc = W[:,64:96]+W[:,96:128] of the 128->32 1x1 weight W
 * (HF = feat - LF folded in).  The channel mix runs between the two box-filter levels on tcgen05 (TF32 operands,
 * fp32 accumulate), so only 64 level-2 box filters remain; guided-filter statistics stay fp32.
 * wpack: 16 KB of TF32-rounded UMMA B tiles [K8 step][16-B chunk][n][4 k]: [Wa;Wb] (n = 64), Wa+Wb (n = 32), Wc (n = 32).
 * stats: paif_gf_guide_stats output.  out: fp32 C4 map, or a bf16 C8 map when out_bf16 != 0.
 * paif_gf_mix_supported: C == 32 and W % 4 == 0 (otherwise use paif_gf_decomp_forward + paif_conv_forward). */
int paif_gf_mix_supported(int C, int H, int W);
int paif_gf_mix_forward(const float* feat, const float* residue, const float* stats, const void* wpack,
                        const float* bias, void* out, int out_bf16, int C, int B, int H, int W, void* stream);
/* The same, also writing mean_a = mean2(A') (fp32 C4 map [B][8][H][W][4]): d out_o / d residue at fixed window
 * statistics, which paif_gf_decomp_backward_saved reads instead of re-running the forward filter. */
int paif_gf_mix_forward_save(const float* feat, const float* residue, const float* stats, const void* wpack,
                             const float* bias, void* out, int out_bf16, float* mean_a,
                             int C, int B, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------
 * bf16 storage mode (forward only; north_star's 1e-2 tier).  Stems, guide and guided-filter
 * statistics stay fp32; from the 1x1 that follows the decomposition (PaifConvDesc.storage = PAIF_STORAGE_F32_BF16)
 * every 32-channel map is a bf16 C8 map [B][4][H][W][8].  All arithmetic is fp32 in registers / TMEM; values are
 * rounded to bf16 (nearest even) once, at the store.
 *   paif_stem_forward_bf16copy  = paif_stem_forward that also writes feat as a bf16 C8 map (the branch residual)
 *   paif_dilconv_forward_bf16   = paif_dilconv_forward   (FFMA kernel) on bf16 maps
 *   paif_spa_fused_forward_bf16 = paif_spa_fused_forward on bf16 maps (no saved attention plane)
 *   paif_eca_apply_bf16         = paif_eca_apply         on bf16 maps (e stays fp32 [B][C])
 *   paif_out_forward_bf16       = paif_out_forward       on a bf16 feature map (fp32 image out)        */
int paif_stem_forward_bf16copy(const float* img, long long stride_b, long long stride_y, long long stride_x,
                               const float* w, const float* slope, float* feat, float* residue, void* feat_bf16,
                               int B, int H, int W, void* stream);
int paif_dilconv_forward_bf16(const void* x, const float* dw, const float* pw, const float* ch_scale,
                              const float* ch_shift, const void* r1, const void* r2, void* out, int add_x,
                              int C, int k, int dil, int B, int H, int W, void* stream);
int paif_spa_fused_forward_bf16(const float* w, int k, const void* ir_f, const void* vis_f, void* agg,
                                int C, int B, int H, int W, void* stream);
int paif_eca_apply_bf16(const void* o, const void* x, const float* e, const float* slope, const void* post_res,
                        void* out, int C, int B, int H, int W, void* stream);
int paif_out_forward_bf16(const void* feat, const float* wm, const float* slope, float* out,
                          int C, int B, int H, int W, void* stream);
/* paif_out_forward with the interior pixels on the tcgen05 engine (TF32 or bf16 operands, fp32 accumulate): the
 * merged 5x5 32->1 stencil is an implicit GEMM with its single output channel padded to 16 accumulator columns.
 * w_mma: interior-class weights packed like a conv weight image with cout = 16 (channel 0 real, the rest zero):
 *   storage F32 : [dx][4][2][dy][16 cout][4 cin] TF32-rounded fp32;  storage BF16: [dx][2][2][dy][16 cout][8 cin] bf16.
 * wm: the 9-class fp32 weights of paif_out_forward (the one-pixel image border is recomputed exactly from them).
 * pre_out may be NULL.  storage: PAIF_STORAGE_F32 (feat fp32 C4) or PAIF_STORAGE_BF16 (feat bf16 C8). */
int paif_out_forward_tc(const void* feat, const void* w_mma, const float* wm, const float* slope,
                        float* out, float* pre_out, int storage, int C, int B, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------
 * Generic dense "same"-padded stride-1 convolution over 1..3 concatenated C4 source maps
 * (torch.cat along channels is a K-loop over sources) with a fused epilogue.  Replaces
 * BasicConv / conv3x3 / nn.Conv2d call sites of ResidualDenseBlock (operations_m.py:435-449),
 * ECABasicBlock (:368-393), ResidualModule (:451-464), Cell_Decom.conv1x1_{lf,hf}
 * (core/model_fusion_auto.py:501-502) and their dgrad (weights transposed + flipped).
 *
 * epilogue, per output element v = accumulator:
 *   v = v*ch_scale[c] + ch_shift[c]            (bias / eval-mode BatchNorm; NULL = skip)
 *   v += pre_res[0] + pre_res[1]               (gradient accumulation before a mask)
 *   out_pre = v                                (optional, saved for backward)
 *   if mask_src: v *= (mask_src > 0 ? 1 : *mask_slope)      (PReLU' of a saved tensor)
 *   else if slope: v = v > 0 ? v : v * *slope               (PReLU)
 *   v *= post_scale
 *   v += post_res[0] + post_res[1] + post_res[2]            (residual adds)
 *   out = v ; out_act2 = PReLU(v, *slope2) (optional) ;
 *   chan_partials[b][tile][c] = sum over the tile's pixels of v (optional, deterministic)
 * The residual sums are formed in an engine-defined order (the tcgen05 engine fetches all maps of a row first and adds
 * them afterwards); with bf16 storage the tcgen05 engine takes at most 4 residual maps per launch (pre + post).
 */
typedef struct PaifConvDesc {
    int B, H, W;
    int nsrc;                 /* 1..3 sources                                      */
    int cin_per_src;          /* channels per source map (multiple of 4; 32)       */
    int cout;                 /* 32 (16 also accepted by the direct engine)        */
    int kh, kw, dil;          /* odd kernel, padding = dil*(k-1)/2                 */
    int engine;               /* PAIF_ENGINE_*                                     */
    const void*  src[3];      /* maps: fp32 C4 or bf16 C8, see `storage`           */
    const float* weight;      /* direct engine: [nsrc][kh*kw][cin_per_src][cout]   */
    const void*  weight_mma;  /* tcgen05 engine: UMMA-packed image, see DESIGN.md  */
    const float* ch_scale;
    const float* ch_shift;
    const void*  pre_res[2];
    void*        out_pre;
    const void*  mask_src;
    const float* mask_slope;
    const float* slope;
    float        post_scale;
    const void*  post_res[3];
    void*        out;
    void*        out_act2;
    const float* slope2;
    float*       chan_partials;
    int          storage;     /* PAIF_STORAGE_* (0 = fp32 everywhere, the default)  */
} PaifConvDesc;

int paif_conv_forward(const PaifConvDesc* desc, void* stream);
/* tcgen05 engine weight image (PaifConvDesc.weight_mma): fp32 values rounded to TF32 (round to nearest),
 * [K-group of KQ channel quads][dx][KQ/2][2 (16-byte chunk)][dy][32 cout][4 cin] — the k row taps of one
 * (dx, 8 input channels) form one UMMA B tile of N = 32*k rows; KQ = paif_conv_tc_kq()
 * (8, or 4 when the weights must be split into passes; 0 = shape not supported by the engine). */
int paif_conv_tc_kq(int nsrc, int k, int dil);
/* the same for PAIF_STORAGE_BF16: bf16 values, [K-group of KQ 8-channel planes][dx][KQ/2][2][dy][32 cout][8 cin];
 * KQ = 4 (a whole 32-channel source), 0 = not supported. */
int paif_conv_tc_kq_bf16(int nsrc, int k, int dil);
/* Launch shape of the tcgen05 engine.  1 (default): persistent — one CTA per SM; the launch's image rows are split into
 * equal bands and the CTAs of a band walk the same rows, one 128-pixel column strip each (no partial last wave, one
 * prologue per SM, neighbouring strips share their border in L2); 0: one CTA per (strip, row chunk, image) tile.
 * Results are bit-identical in both shapes; launches that produce chan_partials are always tiled (their sums are per
 * fixed tile).  Returns the previous setting.  A diagnostic / A-B switch, not part of the reference's API. */
int paif_conv_set_persistent(int on);
/* number of per-image tiles the chosen engine writes into chan_partials ([B][tiles][cout]) */
int paif_conv_num_tiles(int H, int W, int engine);

/* ------------------------------------------------------------------------------------
 * Depthwise k x k dilated convolution on a C4 map — the groups=C BasicConv inside DilConv
 * (operations_m.py:499) and its transpose.  out = dw(relu_in ? max(x,0) : x);
 * then optional  out *= (mask_src > 0)  and  out += post_res.   w: [C][k*k]. */
int paif_dwconv_forward(const float* x, const float* w, int relu_in,
                        const float* mask_src, const float* post_res, float* out,
                        int C, int k, int dil, int B, int H, int W, void* stream);

/* Fused DilConv (operations_m.py:494-506), C = 32, BatchNorm in eval mode folded to scale/shift:
 *   out = ch_scale * pw1x1(dw_kxk_dil(relu(x))) + ch_shift + (add_x ? x : 0) + r1 + r2
 * (add_x = 0 gives one half of SepConv, operations_m.py:509-526; r1, r2 optional residual maps:
 * the Cell_Chain residual, core/model_fusion_auto.py:445, and Cell_Decom's "+ feature", :516).
 * dw: [C][k*k] depthwise taps, pw: [C_out][C_in] pointwise weights.
 * engine: PAIF_ENGINE_TCGEN05 runs the 1x1 on tcgen05 (TF32 operands, fp32 accumulate; measured slower than the
 * all-fp32 FFMA kernel that AUTO and DIRECT use). */
int paif_dilconv_forward(const float* x, const float* dw, const float* pw, const float* ch_scale,
                         const float* ch_shift, const float* r1, const float* r2, float* out,
                         int add_x, int engine, int C, int k, int dil, int B, int H, int W, void* stream);

/* 2-arg ChannelPool — core/model_fusion_auto.py:1352-1355.
 * pooled[B][H][W][4] = (max_c ir, mean_c ir, max_c vis, mean_c vis). */
int paif_channel_pool(const float* ir_f, const float* vis_f, float* pooled,
                      int C, int B, int H, int W, void* stream);

/* spatial_attn_layer_M (BasicConv(4,1,k) + sigmoid) and the blend
 * scale*ir + (1-scale)*vis — core/model_fusion_auto.py:1358-1368, :631-632.
 * w: [4][k*k].  scale_out: [B][H][W] (saved for backward, may be NULL). */
int paif_spa_blend_forward(const float* pooled, const float* w, int k,
                           const float* ir_f, const float* vis_f,
                           float* agg, float* scale_out, int C, int B, int H, int W, void* stream);
/* paif_channel_pool + paif_spa_blend_forward in one kernel (the pooled planes stay in shared memory). */
int paif_spa_fused_forward(const float* w, int k, const float* ir_f, const float* vis_f, float* agg,
                           float* scale_out, int C, int B, int H, int W, void* stream);

/* eca_layer (operations_m.py:340-367): reduce per-tile channel sums (fixed order), mean,
 * Conv1d(1,1,k,pad (k-1)/2, bias=False) across channels, sigmoid.  e: [B][C]. */
int paif_eca_scale(const float* chan_partials, int tiles, const float* w1d, int k,
                   float* e, int C, int B, int H, int W, void* stream);
/* ECABasicBlock tail: out = PReLU(o * e[b][c] + x) (+ post_res) — operations_m.py:389-393 */
int paif_eca_apply(const float* o, const float* x, const float* e, const float* slope,
                   const float* post_res, float* out, int C, int B, int H, int W, void* stream);

/* stem_out (Conv 32->16 3x3, Conv 16->1 3x3, PReLU) + tanh — core/model_fusion_auto.py:615-620,634.
 * The two bias-free convolutions are merged into one 5x5 32->1 stencil; zero padding of the
 * 16-channel intermediate is honoured exactly through 9 border-class weight sets
 * wm[3][3][25][C] (class = (y==0?0 : y==H-1?2 : 1, same for x)).
 * pre_out (optional): pre-activation saved for backward.  out: [B][H][W]. */
int paif_out_forward(const float* feat, const float* wm, const float* slope,
                     float* out, float* pre_out, int C, int B, int H, int W, void* stream);

/* ====================================================================================
 * backward-to-input (autograd of the above, attack/attack.py:501); weight gradients are
 * intentionally not produced.
 * ==================================================================================== */

/* adjoint of paif_out_forward: gs = g * (1 - out^2) * PReLU'(pre); gfeat = stencil^T(gs).
 * Optional second output gfeat_masked = gfeat * PReLU'(mask_src) (feeds the next dgrad). */
int paif_out_backward(const float* g, const float* out, const float* pre_out,
                      const float* wm, const float* slope,
                      float* gfeat, const float* mask_src, const float* mask_slope, float* gfeat_masked,
                      int C, int B, int H, int W, void* stream);

/* out = g * PReLU'(mask_src; *mask_slope) * scale   (elementwise on C4 maps) */
int paif_mask_scale(const float* g, const float* mask_src, const float* mask_slope, float scale,
                    float* out, int C, int B, int H, int W, void* stream);

/* out = a + b (+ c)  elementwise over n floats (n % 4 == 0) — residual adds that could not be
 * fused into a producer's epilogue (Cell_Chain.forward, core/model_fusion_auto.py:445). */
int paif_add_maps(const float* a, const float* b, const float* c, float* out, long long n, void* stream);
/* out = PReLU_slope(x + y); pre_out (optional) = x + y — tail of Spatial_BasicBlock (operations_m.py:203-205). */
int paif_add_act(const float* x, const float* y, const float* slope, float* out, float* pre_out,
                 long long n, void* stream);

/* ECA backward, pass 1: w = o*e + x; gw = gu * PReLU'(w); partial sums of gw*o per (b, tile, c). */
int paif_eca_bwd_pass1(const float* gu, const float* o, const float* x, const float* e,
                       const float* slope, float* gw, float* partials,
                       int C, int B, int H, int W, void* stream);
int paif_eca_bwd_tiles(int H, int W);
/* ECA backward, scalar part: ge = sum gw*o; gm = conv1d^T(ge * e(1-e)) / (H*W).  gm: [B][C] */
int paif_eca_bwd_scale(const float* partials, int tiles, const float* e, const float* w1d, int k,
                       float* gm, int C, int B, int H, int W, void* stream);
/* ECA backward, pass 2: go = gw * e[b][c] + gm[b][c] */
int paif_eca_bwd_pass2(const float* gw, const float* e, const float* gm, float* go,
                       int C, int B, int H, int W, void* stream);

/* adjoint of the blend + spatial attention + ChannelPool. */
int paif_spa_blend_backward_pre(const float* gagg, const float* ir_f, const float* vis_f,
                                const float* scale, float* gpre, int C, int B, int H, int W, void* stream);
int paif_spa_blend_backward(const float* gagg, const float* ir_f, const float* vis_f,
                            const float* scale, const float* gpre, const float* w, int k,
                            float* g_ir_f, float* g_vis_f, int C, int B, int H, int W, void* stream);

/* adjoint of paif_gf_decomp_forward w.r.t. feat (source) and residue (guide).
 * stats: the same guide statistics the forward used (paif_gf_guide_stats).
 * glf1/glf2: gradients w.r.t. the two LF maps.  gfeat: C4 map (written).
 * gres_partial: [paif_gf_guide_parts(C)][B][H][W] partial guide gradients, one plane per channel group
 *   (summed in fixed order by the stem backward).
 * work: caller-owned scratch of paif_gf_backward_work_floats(C,B,H,W) floats. */
long long paif_gf_backward_work_floats(int C, int B, int H, int W);
int paif_gf_guide_parts(int C);
int paif_gf_decomp_backward(const float* feat, const float* residue, const float* stats,
                            const float* glf1, const float* glf2,
                            float* gfeat, float* gres_partial, float* work,
                            int C, int B, int H, int W, void* stream);
/* The adjoint of the FUSED forward (paif_gf_mix_forward_save, autograd of core/model_fusion_auto.py:509-535): gx is the
 * gradient w.r.t. the fused kernel's output (fp32 C4 map), glf1 / glf2 its images under the transposed 1x1
 * (Wa^T gx, Wb^T gx), mean_a the map the forward saved.  The direct guide term sum_o gx_o mean2(A'_o) replaces
 * the forward recompute of paif_gf_decomp_backward (two marching passes + one pointwise pass instead of three). */
int paif_gf_decomp_backward_saved(const float* feat, const float* residue, const float* stats,
                                  const float* glf1, const float* glf2, const float* gx, const float* mean_a,
                                  float* gfeat, float* gres_partial, float* work,
                                  int C, int B, int H, int W, void* stream);

/* stem backward, pass 1: total = sum of up to 4 gradient maps + route(sum_q gres_partial) to the
 * arg-max / arg-min channel of feat (first index on ties); gpre = total * PReLU'(feat). */
int paif_stem_backward_pre(const float* feat, const float* slope,
                           const float* g0, const float* g1, const float* g2, const float* g3,
                           const float* gres_partial, int nparts, float* gpre,
                           int C, int B, int H, int W, void* stream);
/* stem backward, pass 2: gimg = conv3x3^T(gpre) (32 -> 1); gimg: contiguous [B][H][W]. */
int paif_stem_backward(const float* gpre, const float* w, float* gimg,
                       int C, int B, int H, int W, void* stream);

/* ====================================================================================
 * Colour / normalisation glue of the task wrappers: what Network_MM_CompModel.forward does between
 * the fusion net and the segmentation consumer (core/model_fusion_auto.py:712-728 with RGB2YCrCb :69-92 and YCrCb2RGB
 * :94-111) — Cr/Cb of the visible image, YCrCb -> RGB of [fused, Cr, Cb], clamp to [0,1], min-max stretch,
 * x255, (x - mean) / std — as two passes instead of ~15 elementwise launches, and its adjoint as two passes.
 *   fused: [B][H][W]; vis: contiguous RGB [B][3][H][W]; x, gx, gvis: [B][3][H][W]; gfused: [B][H][W]
 *   per_sample != 0: min / max per sample (= the reference at batch 1 applied to every sample);
 *   per_sample == 0: min / max over the whole batch, exactly as the reference wrapper.
 *   Caller-owned scratch with nblk = paif_glue_blocks(H, W): partial [B][nblk][2] float, ties [B][nblk][2] int,
 *   lohi [B][2] float (all three written by the forward and read by the backward), sums [B][nblk][2] float.
 *   mean3 / std3 are HOST pointers to 3 floats.
 * The backward follows autograd of the reference expressions: torch.where clamps pass no gradient where they clamp,
 * torch.min / torch.max share their gradient evenly among tied elements.  gvis holds only the Cr / Cb path; the
 * gradient through the fusion net's Y input is added by the caller (paif_stem_forward_rgb's adjoint). */
int paif_glue_blocks(int H, int W);
int paif_glue_forward(const float* fused, const float* vis, const float* mean3, const float* std3,
                      float* x, float* partial, int* ties, float* lohi, int per_sample,
                      int B, int H, int W, void* stream);
int paif_glue_backward(const float* fused, const float* vis, const float* gx, const float* std3,
                       const float* lohi, const int* ties, float* sums, float* gfused, float* gvis,
                       int per_sample, int B, int H, int W, void* stream);

/* ====================================================================================
 * Whole-network entry point: Network_Fusion_Searched.forward(ir, vis) (core/model_fusion_auto.py:625-635) for the
 * shipped `fusion_at` genotype (test_original.py:711-713 == robust_test.py:255-257), C = 32, eval mode, as ONE call
 * over a caller-owned workspace.  It launches the same kernels in the same order as the per-operator entry points
 * driven by paif_b200/fusion.py (tcgen05 engine), so results are bit-identical; no allocation, no synchronisation.
 *   weights : device pointers to the packed parameters (what paif_b200/fusion.py::_packed builds once per parameter
 *             version: TF32 / bf16 UMMA weight images, folded BatchNorm scale/shift, the fused-decomposition mix tiles,
 *             the merged stem_out stencil); every conv carries its direct-engine image too (may be NULL here).
 *   ir, vis : fp32 1-channel images with arbitrary batch / row / pixel strides (vis is the channel-last strided Y view
 *             the reference wrappers pass, core/model_fusion_auto.py:82-91); out: contiguous [B][1][H][W] fp32.
 *   storage : PAIF_STORAGE_F32 (TF32 MMAs, max-abs 1e-3 tier) or PAIF_STORAGE_BF16 (bf16 maps, 1e-2 tier).
 *   workspace: >= paif_fusion_workspace_bytes(B, H, W, storage) bytes, 256-byte aligned, contents undefined on entry.
 * Needs W % 4 == 0 (paif_gf_mix_supported); other shapes / genotypes go through the per-operator entry points. */
typedef struct PaifFusionConv { const float* direct; const void* mma_tf32; const void* mma_bf16; } PaifFusionConv;
typedef struct PaifFusionRDB { PaifFusionConv conv[3]; const float* slope; } PaifFusionRDB;
typedef struct PaifFusionWeights {
    const float* stem_w[2];          /* stem_1 / stem_2 conv weights [32][9]                         */
    const float* stem_a[2];          /* their PReLU slopes                                            */
    const void*  gfmix_w[2];         /* paif_gf_mix_forward weight images (conv1x1_lf / conv1x1_hf)   */
    const float* c1x1_b[2];          /* conv1x1_lf / conv1x1_hf bias                                  */
    PaifFusionRDB rdb[3];            /* decompation.chain._ops.0, decompation.chain2._ops.0, ._ops.1  */
    PaifFusionConv dil_dense;        /* DilConv as one dense 3x3 (dil 2) convolution: pw[co][ci] dw[ci][t] */
    const float* dil_scale;          /* its folded BatchNorm                                          */
    const float* dil_shift;
    const float* spa_w;              /* spa.spatial.conv.weight [4][k*k]                              */
    int spa_k;
    PaifFusionConv eca_conv1, eca_conv2;
    const float* eca_w1d;            /* eca_layer Conv1d weight [3]                                   */
    const float* eca_a;              /* ECABasicBlock PReLU slope                                     */
    PaifFusionConv res_conv7, res_merged;    /* ResidualModule: 7x7, and 3x3(dil 2) merged with the 1x1 */
    const float* res_scale;          /* its folded BatchNorm                                          */
    const float* res_shift;
    const float* res_a;
    const void*  out_mma_tf32;       /* merged stem_out stencil (paif_out_forward_tc)                 */
    const void*  out_mma_bf16;
    const float* out_wm;
    const float* out_a;
} PaifFusionWeights;
long long paif_fusion_workspace_bytes(int B, int H, int W, int storage);

/* The PGD inner step (attack/attack.py:444-501: forward, loss, backward to the INPUTS) of the same genotype as two calls
 * over one caller-owned workspace, fp32 storage: paif_fusion_forward_save leaves the activations the backward needs at
 * fixed offsets of the workspace (~19 maps of B*H*W*128 bytes, two of them the mean2(A') maps of the fused decomposition), paif_fusion_backward_input turns d loss / d out
 * (gout: [B][1][H][W]) into d loss / d ir and d loss / d vis (contiguous [B][H][W] planes; weight gradients are not
 * produced — nothing in the reference reads them).  Same kernels in the same order as the autograd node of
 * paif_b200/fusion.py, so the gradients are bit-identical to it.  grad_weights: the dgrad images of every convolution
 * (weights flipped and transposed, one PaifFusionConv per group of 32 input channels; BatchNorm scales folded).
 * workspace: >= paif_fusion_train_workspace_bytes(B, H, W) bytes, 256-byte aligned; it must be left untouched between
 * the two calls. */
typedef struct PaifFusionRDBGrad { PaifFusionConv c3[3]; PaifFusionConv c2[2]; PaifFusionConv c1; } PaifFusionRDBGrad;
typedef struct PaifFusionGradWeights {
    PaifFusionRDBGrad rdb[3];            /* same order as PaifFusionWeights.rdb                                  */
    PaifFusionConv dil_dense_d;          /* DilConv's dense form (BatchNorm scale folded)                        */
    PaifFusionConv eca_conv1_d, eca_conv2_d;
    PaifFusionConv res_conv7_d, res_merged_d;
    PaifFusionConv c1x1_d[2][3];         /* folded decomposition 1x1 (lf / hf): groups LF_1e-3, LF_1e-4, z       */
} PaifFusionGradWeights;
long long paif_fusion_train_workspace_bytes(int B, int H, int W);
int paif_fusion_forward_save(const PaifFusionWeights* weights,
                             const float* ir, long long ir_stride_b, long long ir_stride_y, long long ir_stride_x,
                             const float* vis, long long vis_stride_b, long long vis_stride_y, long long vis_stride_x,
                             float* out, void* workspace, long long workspace_bytes, int B, int H, int W, void* stream);
int paif_fusion_backward_input(const PaifFusionWeights* weights, const PaifFusionGradWeights* grad_weights,
                               const float* gout, float* g_ir, float* g_vis,
                               void* workspace, long long workspace_bytes, int B, int H, int W, void* stream);
int paif_fusion_forward(const PaifFusionWeights* weights,
                        const float* ir, long long ir_stride_b, long long ir_stride_y, long long ir_stride_x,
                        const float* vis, long long vis_stride_b, long long vis_stride_y, long long vis_stride_x,
                        float* out, void* workspace, long long workspace_bytes, int storage,
                        int B, int H, int W, void* stream);

/* bf16 C8 map [B][C/8][H][W][8] -> fp32 C4 map [B][C/4][H][W][4].  The backward-to-input chain keeps fp32 gradient
 * maps; this is how it reads the activations a bf16-storage forward saved (net.storage = 'bf16' with requires_grad). */
int paif_widen_bf16_map(const void* src, float* dst, int C, int B, int H, int W, void* stream);

/* ====================================================================================
 * Loss head of the attack (attack/attack.py:446-448 with Seg_loss :103-114):
 *   up = F.interpolate(seg, size=(H, W), mode='bilinear', align_corners=False); CrossEntropyLoss(ignore_index)
 * forward: partial[B][paif_glue_blocks(H, W)] = per-block sums of the per-pixel losses times inv_norm (the caller adds
 *          them up: inv_norm = 1 / #valid gives the reference's mean); gup (optional) [B][K][H][W] receives the gradient
 *          of the loss w.r.t. the UP-SAMPLED logits.
 * backward: gseg[B][K][h][w] = *gscale (device scalar, may be null = 1) x the bilinear adjoint of gup, as a
 *          fixed-order gather (deterministic; the stock backward scatters with float atomics). */
int paif_segloss_forward(const float* seg, const long long* label, float* partial, float* gup,
                         long long ignore_index, float inv_norm, int K, int B, int h, int w, int H, int W, void* stream);
int paif_segloss_backward(const float* gup, const float* gscale, float* gseg,
                          int K, int B, int h, int w, int H, int W, void* stream);

/* ====================================================================================
 * PGD step (attack/attack.py:504-512), in place and for one modality:
 *   delta <- clamp(clamp(delta + alpha * sign(grad), -eps, eps), 0 - x, 1 - x)
 * grad is delta.grad as autograd accumulated it (the reference never zeroes it).  n elements, contiguous fp32. */
int paif_pgd_step(float* delta, const float* grad, const float* x, float alpha, float eps,
                  long long n, void* stream);

/* ====================================================================================
 * evaluation metric: 9x9 confusion matrix (sklearn.metrics.confusion_matrix(labels=0..n-1),
 * robust_test.py:207-211) accumulated on the GPU as int64; the caller all-reduces it (NCCL).
 * conf: [n][n] int64, rows = label, cols = prediction; labels/preds outside 0..n-1 ignored. */
int paif_confusion_accumulate(const long long* label, const long long* pred, long long count,
                              int num_classes, long long* conf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAIF_B200_H */
