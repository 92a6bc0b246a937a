"""GPU: each CUDA kernel, called through the C ABI, against the CPU oracle / plain torch fp32 ops
on small seeded inputs (ragged sizes on purpose: not multiples of any tile)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import paif_b200
from oracle import fusion_oracle as fo
from paif_b200 import _lib, fusion

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_c4(t):
    B, C, H, W = t.shape
    return t.reshape(B, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()


def from_c4(t):
    B, Q, H, W, _ = t.shape
    return t.permute(0, 1, 4, 2, 3).reshape(B, Q * 4, H, W)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rt(B, H, W, engine=_lib.ENGINE_DIRECT, save=False):
    return fusion._Runtime(B, H, W, 32, torch.device(DEV), engine, save)


@pytest.mark.parametrize("shape", [(2, 40, 56), (1, 33, 47), (1, 11, 130)])
def test_stem_and_residue(shape):
    B, H, W = shape
    torch.manual_seed(0)
    img = torch.rand(B, 3, H, W)
    w = torch.randn(32, 1, 3, 3) * 0.3
    a = torch.tensor([0.2])
    ref = F.prelu(F.conv2d(img[:, 0:1], w, None, 1, 1), a)
    ref_res = fo.get_residue(ref)
    v = img.to(DEV).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)[:, 0:1]   # strided view
    feat = torch.empty(B, 8, H, W, 4, device=DEV)
    res = torch.empty(B, H, W, device=DEV)
    wd, ad = w.to(DEV).reshape(32, 9).contiguous(), a.to(DEV)      # keep alive: raw pointers below
    _lib.call("paif_stem_forward", v.data_ptr(), v.stride(0), v.stride(2), v.stride(3),
              wd.data_ptr(), ad.data_ptr(), feat.data_ptr(), res.data_ptr(), B, H, W, stream())
    assert (from_c4(feat).cpu() - ref).abs().max().item() < 1e-5
    assert (res.cpu() - ref_res[:, 0]).abs().max().item() < 1e-5


@pytest.mark.parametrize("shape,smooth", [((2, 40, 56), False), ((1, 33, 47), True), ((1, 10, 70), False),
                                          ((1, 200, 236), False), ((1, 480, 640), True), ((1, 19, 10), False)])
def test_guided_filter_decomposition(shape, smooth):
    B, H, W = shape
    torch.manual_seed(1)
    z = torch.rand(B, 32, H, W)
    if smooth:
        z = F.avg_pool2d(z, 9, 1, 4)
    res = fo.get_residue(z)
    LF, _ = fo.decomposition(z)                      # fp32 oracle (cumsum box filter, as the reference)
    LF64, _ = fo.decomposition(z.double())           # same algorithm in fp64: the exact answer
    zc = to_c4(z).to(DEV)
    lf1, lf2 = torch.empty_like(zc), torch.empty_like(zc)
    resd = res[:, 0].contiguous().to(DEV)
    stats = torch.empty(3, B, H, W, device=DEV)
    _lib.call("paif_gf_guide_stats", resd.data_ptr(), stats.data_ptr(), B, H, W, stream())
    _lib.call("paif_gf_decomp_forward", zc.data_ptr(), resd.data_ptr(), stats.data_ptr(),
              lf1.data_ptr(), lf2.data_ptr(), 32, B, H, W, stream())
    got = torch.cat([from_c4(lf1), from_c4(lf2)], 1).cpu()
    e32 = (got - LF).abs().max().item()
    e64 = (got.double() - LF64).abs().max().item()
    ref_noise = (LF.double() - LF64).abs().max().item()          # the fp32 reference's own rounding error
    # the fp32 cumsum reference is itself 2e-5 .. 5e-4 away from the exact answer (grows with image size);
    # the kernel must be close to the exact answer and within the reference's own noise of the reference
    assert e64 < max(5e-5, 0.5 * ref_noise), (e64, ref_noise)
    assert e32 < 5e-5 + 1.5 * ref_noise, (e32, ref_noise)


def _gf_mix(z, w, bias, out_bf16=False):
    """paif_gf_guide_stats + paif_gf_mix_forward through the C ABI on a [B,32,H,W] CPU feature map."""
    B, _, H, W = z.shape
    zc = to_c4(z).to(DEV)
    resd = fo.get_residue(z)[:, 0].contiguous().to(DEV)
    stats = torch.empty(3, B, H, W, device=DEV)
    wp, bd = fusion._pack_gf_mix(w.to(DEV)), bias.to(DEV)
    _lib.call("paif_gf_guide_stats", resd.data_ptr(), stats.data_ptr(), B, H, W, stream())
    if out_bf16:
        out = torch.empty(B, 4, H, W, 8, device=DEV, dtype=torch.bfloat16)
    else:
        out = torch.empty_like(zc)
    _lib.call("paif_gf_mix_forward", zc.data_ptr(), resd.data_ptr(), stats.data_ptr(), wp.data_ptr(), bd.data_ptr(),
              out.data_ptr(), int(out_bf16), 32, B, H, W, stream())
    torch.cuda.synchronize()
    if out_bf16:
        return out.float().permute(0, 1, 4, 2, 3).reshape(B, 32, H, W).cpu()
    return from_c4(out).cpu()


@pytest.mark.parametrize("shape,smooth", [((2, 40, 56), False), ((1, 33, 48), True), ((1, 10, 72), False),
                                          ((1, 200, 236), False), ((1, 480, 640), True), ((3, 19, 12), False),
                                          ((1, 131, 100), False)])
def test_fused_decomposition_and_1x1(shape, smooth):
    """paif_gf_mix_forward == conv1x1(cat[LF, HF]) of the reference (core/model_fusion_auto.py:509-535), evaluated in
    fp64 by the oracle.  The channel mix runs on TF32 tensor cores: the gate is the TF32 operand rounding (2^-11) of
    the mixed terms, measured against sum_c |W||term|; widths that are not multiples of the 48-column strip, heights
    that force several row chunks per strip and chunk boundaries inside an image are all in the list."""
    B, H, W = shape
    torch.manual_seed(4)
    z = torch.rand(B, 32, H, W)
    if smooth:
        z = F.avg_pool2d(z, 9, 1, 4)
    w = torch.randn(32, 128, 1, 1) * 0.15
    bias = torch.randn(32) * 0.1
    LF64, HF64 = fo.decomposition(z.double())
    cat = torch.cat([LF64, HF64], 1)
    ref = F.conv2d(cat, w.double(), bias.double())
    scale = F.conv2d(cat.abs(), w.double().abs()) + 1.0
    got = _gf_mix(z, w, bias)
    err = ((got.double() - ref).abs() / scale).max().item()
    assert err < 1.5e-3, err                                     # ~3 TF32 roundings of O(1) terms
    assert (got.double() - ref).abs().max().item() < 5e-3
    got16 = _gf_mix(z, w, bias, out_bf16=True)
    assert ((got16.double() - ref).abs() / scale).max().item() < 1.5e-3 + 2.0 ** -8


def test_fused_decomposition_is_deterministic_and_batch_position_invariant():
    torch.manual_seed(5)
    z = torch.rand(3, 32, 70, 100)
    w, bias = torch.randn(32, 128, 1, 1) * 0.15, torch.randn(32) * 0.1
    a = _gf_mix(z, w, bias)
    assert torch.equal(a, _gf_mix(z, w, bias))
    # the chunk grid depends on (batch x strips, H): equal-shaped launches give every image the same bits
    zz = torch.cat([z[2:3], z[0:1], z[1:2]])
    assert torch.equal(a[2:3], _gf_mix(zz, w, bias)[0:1])


@pytest.mark.parametrize("k,dil,nsrc", [(3, 1, 1), (3, 1, 3), (3, 2, 1), (7, 1, 1), (1, 1, 3), (5, 2, 2)])
def test_conv_direct_with_epilogue(k, dil, nsrc):
    B, H, W = 2, 21, 139
    torch.manual_seed(2)
    xs = [torch.randn(B, 32, H, W) for _ in range(nsrc)]
    w = torch.randn(32, 32 * nsrc, k, k) * 0.1
    cs, sh = torch.rand(32) + 0.5, torch.randn(32) * 0.1
    a = torch.tensor([0.3])
    r1, r2 = torch.randn(B, 32, H, W), torch.randn(B, 32, H, W)
    pad = dil * (k - 1) // 2
    pre = F.conv2d(torch.cat(xs, 1), w, None, 1, pad, dil) * cs.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref = F.prelu(pre, a) * 0.5 + r1 + r2
    r = rt(B, H, W, save=True)
    cw = fusion._ConvW(w.to(DEV), nsrc, k, dil)
    out, opre, act2, parts = r.conv([to_c4(x).to(DEV) for x in xs], cw, ch_scale=cs.to(DEV), ch_shift=sh.to(DEV),
                                    slope=a.to(DEV), post_scale=0.5, post_res=[to_c4(r1).to(DEV), to_c4(r2).to(DEV)],
                                    want_pre=True, act2_slope=a.to(DEV), want_partials=True)
    tol = 2e-4 if k == 7 else 5e-5
    assert (from_c4(out).cpu() - ref).abs().max().item() < tol
    assert (from_c4(opre).cpu() - pre).abs().max().item() < tol
    assert (from_c4(act2).cpu() - F.prelu(ref, a)).abs().max().item() < tol
    sums = parts.sum(1).cpu()
    assert (sums - ref.sum((2, 3))).abs().max().item() < 2e-2


def test_conv_direct_backward_style_epilogue():
    B, H, W = 1, 17, 40
    torch.manual_seed(3)
    x = torch.randn(B, 32, H, W)
    w = torch.randn(32, 32, 3, 3) * 0.1
    p1, m = torch.randn(B, 32, H, W), torch.randn(B, 32, H, W)
    a = torch.tensor([0.25])
    ref = (F.conv2d(x, w, None, 1, 1) + p1) * torch.where(m > 0, torch.ones_like(m), a.expand_as(m))
    r = rt(B, H, W)
    out = r.conv([to_c4(x).to(DEV)], fusion._ConvW(w.to(DEV), 1, 3, 1), pre_res=[to_c4(p1).to(DEV)],
                 mask_src=to_c4(m).to(DEV), mask_slope=a.to(DEV))[0]
    assert (from_c4(out).cpu() - ref).abs().max().item() < 5e-5


def test_dwconv_pool_spa_out():
    B, H, W = 2, 19, 45
    torch.manual_seed(4)
    x, y = torch.randn(B, 32, H, W), torch.randn(B, 32, H, W)
    r = rt(B, H, W)
    # depthwise dilated conv with ReLU on the input
    wd = torch.randn(32, 1, 3, 3)
    ref = F.conv2d(F.relu(x), wd, None, 1, 2, 2, groups=32)
    got = r.dwconv(to_c4(x).to(DEV), wd.reshape(32, 9).to(DEV), 3, 2, True)
    assert (from_c4(got).cpu() - ref).abs().max().item() < 1e-5
    # channel pool + spatial attention + blend
    sd = {"spa.spatial.conv.weight": torch.randn(1, 4, 5, 5) * 0.2}
    s = fo.spatial_attn(sd, x, y)
    ref = s * x + (1 - s) * y
    xc, yc = to_c4(x).to(DEV), to_c4(y).to(DEV)
    pooled = torch.empty(B, H, W, 4, device=DEV)
    _lib.call("paif_channel_pool", xc.data_ptr(), yc.data_ptr(), pooled.data_ptr(), 32, B, H, W, stream())
    agg, sc = torch.empty_like(xc), torch.empty(B, H, W, device=DEV)
    wspa = sd["spa.spatial.conv.weight"].reshape(4, 25).to(DEV)
    _lib.call("paif_spa_blend_forward", pooled.data_ptr(), wspa.data_ptr(), 5, xc.data_ptr(), yc.data_ptr(),
              agg.data_ptr(), sc.data_ptr(), 32, B, H, W, stream())
    assert (sc.cpu() - s[:, 0]).abs().max().item() < 1e-5
    assert (from_c4(agg).cpu() - ref).abs().max().item() < 1e-5
    # the same in one kernel (pooled planes stay in shared memory)
    agg2, sc2 = torch.empty_like(xc), torch.empty(B, H, W, device=DEV)
    _lib.call("paif_spa_fused_forward", wspa.data_ptr(), 5, xc.data_ptr(), yc.data_ptr(), agg2.data_ptr(), sc2.data_ptr(),
              32, B, H, W, stream())
    assert (sc2.cpu() - s[:, 0]).abs().max().item() < 1e-5
    assert (from_c4(agg2).cpu() - ref).abs().max().item() < 1e-5
    # merged stem_out + PReLU + tanh
    w1, w2, a = torch.randn(16, 32, 3, 3) * 0.1, torch.randn(1, 16, 3, 3) * 0.1, torch.tensor([0.25])
    ref = torch.tanh(F.prelu(F.conv2d(F.conv2d(x, w1, None, 1, 1), w2, None, 1, 1), a))
    wm = fusion._merge_stem_out(w1, w2).to(DEV)
    out, pre = torch.empty(B, 1, H, W, device=DEV), torch.empty(B, H, W, device=DEV)
    ad = a.to(DEV)
    _lib.call("paif_out_forward", xc.data_ptr(), wm.data_ptr(), ad.data_ptr(), out.data_ptr(),
              pre.data_ptr(), 32, B, H, W, stream())
    assert (out.cpu() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("k,dil,nres", [(3, 2, 2), (3, 1, 0), (3, 2, 1)])
def test_dilconv_fused(k, dil, nres):
    """DilConv (operations_m.py:494-506) + chain residuals in one kernel, against plain torch fp32 ops."""
    B, H, W = 2, 19, 45
    torch.manual_seed(7)
    x = torch.randn(B, 32, H, W)
    dw = torch.randn(32, 1, k, k) * 0.3
    pw = torch.randn(32, 32, 1, 1) * 0.2
    cs, sh = torch.rand(32) + 0.5, torch.randn(32) * 0.1
    res = [torch.randn(B, 32, H, W) for _ in range(nres)]
    pad = dil * (k - 1) // 2
    t = F.conv2d(F.relu(x), dw, None, 1, pad, dil, groups=32)
    ref = F.conv2d(t, pw) * cs.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1) + x + sum(res)
    xc = to_c4(x).to(DEV)
    out = torch.empty_like(xc)
    dwd, pwd = dw.reshape(32, -1).contiguous().to(DEV), pw.reshape(32, 32).contiguous().to(DEV)
    csd, shd = cs.to(DEV), sh.to(DEV)
    rd = [to_c4(r).to(DEV) for r in res]
    _lib.call("paif_dilconv_forward", xc.data_ptr(), dwd.data_ptr(), pwd.data_ptr(), csd.data_ptr(), shd.data_ptr(),
              rd[0].data_ptr() if nres > 0 else None, rd[1].data_ptr() if nres > 1 else None, out.data_ptr(),
              1, _lib.ENGINE_DIRECT, 32, k, dil, B, H, W, stream())
    assert (from_c4(out).cpu() - ref).abs().max().item() < 5e-5
    _lib.call("paif_dilconv_forward", xc.data_ptr(), dwd.data_ptr(), pwd.data_ptr(), csd.data_ptr(), shd.data_ptr(),
              None, None, out.data_ptr(), 0, _lib.ENGINE_DIRECT, 32, k, dil, B, H, W, stream())   # add_x = 0: half a SepConv
    assert (from_c4(out).cpu() - (ref - x - sum(res))).abs().max().item() < 5e-5
    # tensor-core engine: TF32 operands in the 1x1
    _lib.call("paif_dilconv_forward", xc.data_ptr(), dwd.data_ptr(), pwd.data_ptr(), csd.data_ptr(), shd.data_ptr(),
              rd[0].data_ptr() if nres > 0 else None, rd[1].data_ptr() if nres > 1 else None, out.data_ptr(),
              1, _lib.ENGINE_TCGEN05, 32, k, dil, B, H, W, stream())
    assert (from_c4(out).cpu() - ref).abs().max().item() < 2.0 ** -9 * t.abs().max().item() * 4


def test_confusion_matrix_kernel_is_exact():
    g = torch.Generator().manual_seed(5)
    label = torch.randint(0, 10, (3, 97, 131), generator=g)
    label[0, :5] = 255                                          # ignore_index of the loss
    pred = torch.randint(0, 9, (3, 97, 131), generator=g)
    ref = fo.confusion_matrix(label, pred, 9)
    conf = torch.zeros(9, 9, dtype=torch.int64, device=DEV)
    ld, pd = label.to(DEV), pred.to(DEV)
    for _ in range(2):
        _lib.call("paif_confusion_accumulate", ld.data_ptr(), pd.data_ptr(), label.numel(), 9,
                  conf.data_ptr(), stream())
    assert torch.equal(conf.cpu(), 2 * ref)


@pytest.mark.parametrize("B,H,W,k,dil,nsrc", [(1, 20, 128, 3, 1, 1), (2, 37, 200, 3, 1, 1), (1, 40, 300, 3, 1, 3),
                                               (1, 33, 130, 3, 2, 1), (1, 40, 256, 7, 1, 1), (2, 21, 139, 1, 1, 3),
                                               (1, 70, 640, 3, 1, 2), (1, 24, 150, 5, 1, 1)])
def test_conv_tcgen05_matches_direct_engine(B, H, W, k, dil, nsrc):
    """TF32 tensor-core engine vs torch fp64 and vs the exact-fp32 direct engine, with a full epilogue."""
    torch.manual_seed(6)
    xs = [torch.randn(B, 32, H, W) for _ in range(nsrc)]
    w = torch.randn(32, 32 * nsrc, k, k) * 0.1
    r1 = torch.randn(B, 32, H, W)
    a = torch.tensor([0.3])
    pre = F.conv2d(torch.cat(xs, 1).double(), w.double(), None, 1, dil * (k - 1) // 2, dil).float()
    ref = F.prelu(pre, a) * 0.5 + r1
    cw = fusion._ConvW(w.to(DEV), nsrc, k, dil)
    assert cw.mma is not None
    xc, r1c, ad = [to_c4(x).to(DEV) for x in xs], to_c4(r1).to(DEV), a.to(DEV)
    res = {}
    for eng in (_lib.ENGINE_DIRECT, _lib.ENGINE_TCGEN05):
        out, opre, _, parts = rt(B, H, W, eng).conv(xc, cw, slope=ad, post_scale=0.5, post_res=[r1c], want_pre=True,
                                                     want_partials=True)
        res[eng] = (from_c4(out).cpu(), from_c4(opre).cpu(), parts.sum(1).cpu())
    scale = pre.abs().max().item()
    tf32 = 2.0 ** -10                                # TF32 operand truncation bound per product
    assert (res[_lib.ENGINE_TCGEN05][0] - ref).abs().max().item() < 2 * tf32 * scale
    assert (res[_lib.ENGINE_TCGEN05][1] - pre).abs().max().item() < 2 * tf32 * scale
    assert (res[_lib.ENGINE_DIRECT][0] - ref).abs().max().item() < 1e-4 * max(1.0, scale)
    assert (res[_lib.ENGINE_TCGEN05][2] - ref.sum((2, 3))).abs().max().item() < 1e-3 * H * W


@pytest.mark.parametrize("k,dil,nsrc", [(3, 1, 1), (3, 1, 3), (3, 2, 1), (1, 1, 3), (5, 1, 1)])
def test_conv_tcgen05_full_size_tiling(k, dil, nsrc):
    """Enough tiles that the engine uses 32-row chunks: exercises TMEM slot reuse, the slot-ring wrap inside one
    wide-N MMA, ragged last chunk / last strip.  Checked against the exact-fp32 direct engine on the GPU."""
    B, H, W = 37, 67, 507
    torch.manual_seed(8)
    g = torch.Generator(device=DEV).manual_seed(8)
    xs = [torch.randn(B, 8, H, W, 4, device=DEV, generator=g) for _ in range(nsrc)]
    r1 = torch.randn(B, 8, H, W, 4, device=DEV, generator=g)
    m = torch.randn(B, 8, H, W, 4, device=DEV, generator=g)
    w = torch.randn(32, 32 * nsrc, k, k, device=DEV, generator=g) * 0.1
    a = torch.tensor([0.3], device=DEV)
    cw = fusion._ConvW(w, nsrc, k, dil)
    assert cw.mma is not None
    outs = {}
    for eng in (_lib.ENGINE_DIRECT, _lib.ENGINE_TCGEN05):
        o1 = rt(B, H, W, eng).conv(xs, cw, slope=a, post_scale=0.5, post_res=[r1])[0]
        o2 = rt(B, H, W, eng).conv(xs, cw, pre_res=[r1], mask_src=m, mask_slope=a)[0]
        outs[eng] = (o1, o2)
    scale = outs[_lib.ENGINE_DIRECT][0].abs().max().item()
    for i in range(2):
        err = (outs[_lib.ENGINE_TCGEN05][i] - outs[_lib.ENGINE_DIRECT][i]).abs().max().item()
        assert err < 2.0 ** -9 * max(scale, 1.0), (i, err, scale)


@pytest.mark.parametrize("B,H,W,k,dil,nsrc", [(3, 50, 200, 3, 1, 1), (5, 61, 300, 3, 2, 1), (3, 50, 200, 7, 1, 1),
                                               (7, 45, 130, 1, 1, 3), (3, 77, 640, 3, 1, 3), (2, 480, 640, 7, 1, 1),
                                               (3, 77, 300, 3, 2, 1), (5, 53, 520, 5, 1, 1), (4, 480, 640, 3, 2, 1)])
def test_conv_tcgen05_persistent_launch_is_bit_identical_to_tiled(B, H, W, k, dil, nsrc):
    """The persistent launch (one CTA per SM, equal contiguous shares of the B*strips*H output rows, a share crossing
    strip / image boundaries as separate segments) against the tiled launch of the same kernel: same bits (every output
    row accumulates its taps in the same order whatever the partition), and both against the exact-fp32 direct engine."""
    g = torch.Generator(device=DEV).manual_seed(12)
    xs = [torch.randn(B, 8, H, W, 4, device=DEV, generator=g) for _ in range(nsrc)]
    r1 = torch.randn(B, 8, H, W, 4, device=DEV, generator=g)
    w = torch.randn(32, 32 * nsrc, k, k, device=DEV, generator=g) * 0.1
    a = torch.tensor([0.3], device=DEV)
    cw = fusion._ConvW(w, nsrc, k, dil)
    lib = _lib.load()
    outs = {}
    prev = lib.paif_conv_set_persistent(1)
    try:
        for mode in (1, 0):
            lib.paif_conv_set_persistent(mode)
            o, opre, act2, _ = rt(B, H, W, _lib.ENGINE_TCGEN05).conv(xs, cw, slope=a, post_scale=0.5, post_res=[r1],
                                                                      want_pre=True)
            outs[mode] = (o.clone(), opre.clone())
    finally:
        lib.paif_conv_set_persistent(prev)
    assert torch.equal(outs[1][0], outs[0][0]) and torch.equal(outs[1][1], outs[0][1])
    ref = rt(B, H, W, _lib.ENGINE_DIRECT).conv(xs, cw, slope=a, post_scale=0.5, post_res=[r1])[0]
    scale = ref.abs().max().item()
    assert (outs[1][0] - ref).abs().max().item() < 2.0 ** -9 * max(scale, 1.0)


@pytest.mark.parametrize("shape", [(2, 19, 45), (1, 70, 300), (3, 33, 128)])
def test_stem_out_on_tensor_cores_matches_ffma_kernel(shape):
    """paif_out_forward_tc (interior pixels as an implicit GEMM with TF32 / bf16 operands, border pixels exact) against
    the exact-fp32 stencil kernel on the same features."""
    B, H, W = shape
    torch.manual_seed(9)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
    with torch.no_grad():
        net.stem_out[0].weight.mul_(3.0)                    # default init gives tiny outputs; make the test bite
    p = net._packed(False)
    x = torch.randn(B, 32, H, W)
    x4 = to_c4(x).to(DEV)
    o_ref, pre_ref = torch.empty(B, 1, H, W, device=DEV), torch.empty(B, H, W, device=DEV)
    _lib.call("paif_out_forward", x4.data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(), o_ref.data_ptr(),
              pre_ref.data_ptr(), 32, B, H, W, stream())
    o_tc, pre_tc = torch.full_like(o_ref, float("nan")), torch.full_like(pre_ref, float("nan"))
    _lib.call("paif_out_forward_tc", x4.data_ptr(), p["out_mma"].data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(),
              o_tc.data_ptr(), pre_tc.data_ptr(), _lib.STORAGE_F32, 32, B, H, W, stream())
    scale = pre_ref.abs().max().item()
    assert scale > 0.05
    assert (pre_tc - pre_ref).abs().max().item() < 2.0 ** -9 * max(scale, 1.0)
    assert (o_tc - o_ref).abs().max().item() < 2.0 ** -9 * max(scale, 1.0)
    # the one-pixel border is computed by the exact kernel in both paths
    for sl in ((slice(None), 0), (slice(None), H - 1), (slice(None), slice(None), 0), (slice(None), slice(None), W - 1)):
        assert torch.equal(pre_tc[sl], pre_ref[sl])
    # bf16 feature map
    x16 = x.to(torch.bfloat16)
    xc8 = x16.reshape(B, 4, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(DEV)
    x4b = to_c4(x16.float()).to(DEV)
    _lib.call("paif_out_forward", x4b.data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(), o_ref.data_ptr(),
              pre_ref.data_ptr(), 32, B, H, W, stream())
    _lib.call("paif_out_forward_tc", xc8.data_ptr(), p["out_mma16"].data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(),
              o_tc.data_ptr(), None, _lib.STORAGE_BF16, 32, B, H, W, stream())
    assert (o_tc - o_ref).abs().max().item() < 2.0 ** -7 * max(scale, 1.0)          # bf16-rounded weights
