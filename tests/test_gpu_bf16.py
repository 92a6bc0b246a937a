"""GPU: the bf16 storage mode (north_star's "1e-2 in bf16" tier).

bf16 maps are C8 ([B][C/8][H][W][8]); all arithmetic is fp32 (TMEM accumulators / registers) with one
round-to-nearest-even at the store.  Kernel tests therefore compare against torch fp64 on the SAME
bf16-representable inputs and allow only the output rounding (2^-8 relative) plus fp32 summation noise;
the whole-network tests hold the fused image to the 1e-2 gate against the reference golden outputs."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import paif_b200
from oracle import fusion_oracle as fo
from paif_b200 import _lib, fusion
from paif_testutil import GOLDEN_CASES, golden_genotype, load_golden, strided_vis

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_ULP = 2.0 ** -8          # half a unit in the last place of an 8-bit significand, relative


def to_c8(t):
    """[B,C,H,W] fp32 -> bf16 C8 map [B,C/8,H,W,8]."""
    B, C, H, W = t.shape
    return t.reshape(B, C // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def from_c8(t):
    B, P, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(B, P * 8, H, W)


def to_c4(t):
    B, C, H, W = t.shape
    return t.reshape(B, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()


def bf(t):
    """round to bf16 and back: the values a bf16 map can hold."""
    return t.to(torch.bfloat16).float()


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rt16(B, H, W):
    return fusion._Runtime(B, H, W, 32, torch.device(DEV), _lib.ENGINE_TCGEN05, False, bf16=True)


def close_bf16(got, ref, extra=0.0):
    """|got - ref| <= one bf16 rounding of ref (+ fp32 accumulation slack relative to the tensor's scale)."""
    tol = BF16_ULP * ref.abs() + (2e-5 + extra) * ref.abs().max()
    bad = ((got - ref).abs() > tol)
    assert not bad.any(), ((got - ref).abs().max().item(), ref.abs().max().item(), int(bad.sum()))


@pytest.mark.parametrize("B,H,W,k,dil,nsrc", [(1, 20, 128, 3, 1, 1), (2, 37, 200, 3, 1, 3), (1, 33, 130, 3, 2, 1),
                                               (1, 40, 256, 7, 1, 1), (2, 21, 139, 1, 1, 3), (1, 24, 150, 5, 1, 2),
                                               (1, 30, 140, 5, 2, 1), (1, 26, 131, 7, 2, 1)])
def test_conv_bf16_engine_with_epilogue(B, H, W, k, dil, nsrc):
    torch.manual_seed(11)
    xs = [bf(torch.randn(B, 32, H, W)) for _ in range(nsrc)]
    w = bf(torch.randn(32, 32 * nsrc, k, k) * 0.1)
    cs, sh = torch.rand(32) + 0.5, torch.randn(32) * 0.1
    r1, r2 = bf(torch.randn(B, 32, H, W)), bf(torch.randn(B, 32, H, W))
    a = torch.tensor([0.3])
    acc = F.conv2d(torch.cat(xs, 1).double(), w.double(), None, 1, dil * (k - 1) // 2, dil)
    pre = acc * cs.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    ref = F.prelu(pre, a.double()) * 0.5 + r1.double() + r2.double()
    cw = fusion._ConvW(w.to(DEV), nsrc, k, dil)
    assert cw.mma16 is not None and cw.mma16.dtype == torch.bfloat16
    out, opre, act2, parts = rt16(B, H, W).conv(
        [to_c8(x).to(DEV) for x in xs], cw, ch_scale=cs.to(DEV), ch_shift=sh.to(DEV), slope=a.to(DEV), post_scale=0.5,
        post_res=[to_c8(r1).to(DEV), to_c8(r2).to(DEV)], want_pre=True, act2_slope=a.to(DEV), want_partials=True)
    assert out.dtype == torch.bfloat16 and out.shape == (B, 4, H, W, 8)
    slack = 1e-6 * k * k * nsrc                      # fp32 accumulation of exact bf16 x bf16 products
    close_bf16(from_c8(out).cpu().double(), ref, slack)
    close_bf16(from_c8(opre).cpu().double(), pre, slack)
    close_bf16(from_c8(act2).cpu().double(), F.prelu(ref, a.double()), slack + BF16_ULP)   # act2 of the unrounded out
    assert (parts.sum(1).cpu().double() - ref.sum((2, 3))).abs().max().item() < 1e-3 * H * W


def test_conv_bf16_backward_style_epilogue_and_mask():
    B, H, W = 1, 17, 140
    torch.manual_seed(12)
    x, p1, m = (bf(torch.randn(B, 32, H, W)) for _ in range(3))
    w = bf(torch.randn(32, 32, 3, 3) * 0.1)
    a = torch.tensor([0.25])
    ref = (F.conv2d(x.double(), w.double(), None, 1, 1) + p1.double()) * torch.where(m > 0, 1.0, 0.25).double()
    out = rt16(B, H, W).conv([to_c8(x).to(DEV)], fusion._ConvW(w.to(DEV), 1, 3, 1), pre_res=[to_c8(p1).to(DEV)],
                             mask_src=to_c8(m).to(DEV), mask_slope=a.to(DEV))[0]
    close_bf16(from_c8(out).cpu().double(), ref, 1e-5)


def test_conv_fp32_sources_bf16_outputs():
    """storage F32_BF16: the 1x1 after the decomposition (fp32 LF maps + fp32 features in, bf16 map out)."""
    B, H, W = 2, 23, 150
    torch.manual_seed(13)
    xs = [torch.randn(B, 32, H, W) for _ in range(3)]
    w = torch.randn(32, 96, 1, 1) * 0.1
    bias = torch.randn(32) * 0.1
    ref = F.conv2d(torch.cat(xs, 1).double(), w.double(), bias.double())
    out = rt16(B, H, W).conv([to_c4(x).to(DEV) for x in xs], fusion._ConvW(w.to(DEV), 3, 1, 1), ch_shift=bias.to(DEV),
                             src_fp32=True)[0]
    assert out.dtype == torch.bfloat16
    got = from_c8(out).cpu().double()
    assert ((got - ref).abs() <= BF16_ULP * ref.abs() + 2.0 ** -9 * ref.abs().max()).all()     # TF32 operands + bf16 store


def test_conv_bf16_full_size_tiling():
    """32-row chunks, TMEM slot reuse, ragged last chunk / strip, single-pass 7x7 — against the fp32 direct engine on
    the same bf16-representable data."""
    B, H, W = 19, 67, 507
    g = torch.Generator(device=DEV).manual_seed(14)
    for k, dil, nsrc in ((3, 1, 3), (7, 1, 1), (3, 2, 1)):
        xs = [torch.randn(B, 32, H, W, device=DEV, generator=g).to(torch.bfloat16).float() for _ in range(nsrc)]
        r1 = torch.randn(B, 32, H, W, device=DEV, generator=g).to(torch.bfloat16).float()
        w = (torch.randn(32, 32 * nsrc, k, k, device=DEV, generator=g) * 0.1).to(torch.bfloat16).float()
        a = torch.tensor([0.3], device=DEV)
        cw = fusion._ConvW(w, nsrc, k, dil)
        ref = fusion._Runtime(B, H, W, 32, torch.device(DEV), _lib.ENGINE_DIRECT, False).conv(
            [to_c4(x) for x in xs], cw, slope=a, post_scale=0.5, post_res=[to_c4(r1)])[0]
        ref = ref.permute(0, 1, 4, 2, 3).reshape(B, 32, H, W)
        out = rt16(B, H, W).conv([to_c8(x) for x in xs], cw, slope=a, post_scale=0.5, post_res=[to_c8(r1)])[0]
        got = from_c8(out)
        tol = BF16_ULP * ref.abs() + 1e-4 * ref.abs().max()
        assert ((got - ref).abs() <= tol).all(), (k, dil, nsrc, (got - ref).abs().max().item())


@pytest.mark.parametrize("k,dil,nres", [(3, 2, 2), (3, 1, 0)])
def test_dilconv_bf16(k, dil, nres):
    B, H, W = 2, 19, 45
    torch.manual_seed(15)
    x = bf(torch.randn(B, 32, H, W))
    dw, pw = torch.randn(32, 1, k, k) * 0.3, torch.randn(32, 32, 1, 1) * 0.2
    cs, sh = torch.rand(32) + 0.5, torch.randn(32) * 0.1
    res = [bf(torch.randn(B, 32, H, W)) for _ in range(nres)]
    t = F.conv2d(F.relu(x).double(), dw.double(), None, 1, dil * (k - 1) // 2, dil, groups=32)
    ref = F.conv2d(t, pw.double()) * cs.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1) + x.double()
    for r in res:
        ref = ref + r.double()
    xc = to_c8(x).to(DEV)
    out = torch.empty_like(xc)
    rd = [to_c8(r).to(DEV) for r in res]
    dwd, pwd = dw.reshape(32, -1).contiguous().to(DEV), pw.reshape(32, 32).contiguous().to(DEV)
    csd, shd = cs.to(DEV), sh.to(DEV)
    _lib.call("paif_dilconv_forward_bf16", xc.data_ptr(), dwd.data_ptr(), pwd.data_ptr(), csd.data_ptr(), shd.data_ptr(),
              rd[0].data_ptr() if nres > 0 else None, rd[1].data_ptr() if nres > 1 else None, out.data_ptr(),
              1, 32, k, dil, B, H, W, stream())
    close_bf16(from_c8(out).cpu().double(), ref, 1e-5)


def test_spa_eca_out_bf16():
    B, H, W = 2, 19, 45
    torch.manual_seed(16)
    x, y = bf(torch.randn(B, 32, H, W)), bf(torch.randn(B, 32, H, W))
    # channel pool + spatial attention + blend
    sd = {"spa.spatial.conv.weight": (torch.randn(1, 4, 5, 5) * 0.2).double()}
    s = fo.spatial_attn(sd, x.double(), y.double())
    ref = s * x.double() + (1 - s) * y.double()
    xc, yc = to_c8(x).to(DEV), to_c8(y).to(DEV)
    agg = torch.empty_like(xc)
    w4 = sd["spa.spatial.conv.weight"].float().reshape(4, -1).contiguous().to(DEV)
    _lib.call("paif_spa_fused_forward_bf16", w4.data_ptr(), 5, xc.data_ptr(), yc.data_ptr(), agg.data_ptr(),
              32, B, H, W, stream())
    close_bf16(from_c8(agg).cpu().double(), ref, 1e-5)
    # ECA apply: PReLU(o * e + x) + r
    e = torch.rand(B, 32)
    a = torch.tensor([0.2])
    r = bf(torch.randn(B, 32, H, W))
    ref = F.prelu(x.double() * e.double().view(B, 32, 1, 1) + y.double(), a.double()) + r.double()
    out = torch.empty_like(xc)
    ed, ad, rc = e.to(DEV), a.to(DEV), to_c8(r).to(DEV)
    _lib.call("paif_eca_apply_bf16", xc.data_ptr(), yc.data_ptr(), ed.data_ptr(), ad.data_ptr(), rc.data_ptr(),
              out.data_ptr(), 32, B, H, W, stream())
    close_bf16(from_c8(out).cpu().double(), ref, 1e-5)
    # stem_out + tanh on a bf16 feature map == the fp32 kernel on the same (bf16-representable) values
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
    p = net._packed(False)
    o16, o32 = torch.empty(B, 1, H, W, device=DEV), torch.empty(B, 1, H, W, device=DEV)
    x4 = to_c4(x).to(DEV)
    _lib.call("paif_out_forward_bf16", xc.data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(), o16.data_ptr(),
              32, B, H, W, stream())
    _lib.call("paif_out_forward", x4.data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(), o32.data_ptr(), None,
              32, B, H, W, stream())
    assert torch.equal(o16, o32)


def test_stem_bf16_copy():
    B, H, W = 2, 21, 70
    torch.manual_seed(17)
    img = torch.rand(B, 1, H, W, device=DEV)
    w, a = torch.randn(32, 9, device=DEV) * 0.3, torch.tensor([0.25], device=DEV)
    f = torch.empty(B, 8, H, W, 4, device=DEV)
    f2 = torch.empty_like(f)
    g, g2 = torch.empty(B, H, W, device=DEV), torch.empty(B, H, W, device=DEV)
    f16 = torch.empty(B, 4, H, W, 8, device=DEV, dtype=torch.bfloat16)
    args = (img.data_ptr(), img.stride(0), img.stride(2), img.stride(3), w.data_ptr(), a.data_ptr())
    _lib.call("paif_stem_forward", *args, f.data_ptr(), g.data_ptr(), B, H, W, stream())
    _lib.call("paif_stem_forward_bf16copy", *args, f2.data_ptr(), g2.data_ptr(), f16.data_ptr(), B, H, W, stream())
    assert torch.equal(f, f2) and torch.equal(g, g2)
    want = f.permute(0, 1, 4, 2, 3).reshape(B, 32, H, W).to(torch.bfloat16)
    assert torch.equal(from_c8(f16), want.float())


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_bf16_storage_matches_reference_golden(case):
    """north_star: fused image within max-abs 1e-2 of the reference's fp32 PyTorch model in bf16 mode."""
    g = load_golden(case)
    net = paif_b200.Network_Fusion_Searched(32, None, golden_genotype(g))
    net.load_state_dict(g["state_dict"], strict=True)
    net = net.to(DEV).eval()
    net.storage = 'bf16'
    with torch.no_grad():
        try:
            out = net(g["ir"].to(DEV), strided_vis(g["vis"].to(DEV)))
        except NotImplementedError:
            pytest.skip("this fixture's genotype uses a primitive without a bf16 forward (fp32 storage only)")
    err = (out.cpu() - g["out"]).abs().max().item()
    assert err < 1e-2, (case, err)


def test_bf16_storage_full_size_and_mode_rules():
    """480x640: bf16 mode against the fp32 CPU oracle (1e-2 gate) and against the module's own fp32 mode; the mode is
    refuses the exact-fp32 engine."""
    g = load_golden("seed1_random_1x48x72")
    sd = g["state_dict"]
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    gen = torch.Generator().manual_seed(22)
    ir, vis = torch.rand(2, 1, 480, 640, generator=gen), torch.rand(2, 3, 480, 640, generator=gen)
    ref = fo.fusion_forward(sd, paif_b200.fusion_at, ir[:1], vis[:1])
    with torch.no_grad():
        out32 = net(ir.to(DEV), vis.to(DEV))
        net.storage = 'bf16'
        out16 = net(ir.to(DEV), vis.to(DEV))
        launches16 = net.last_launches
        again = net(ir.to(DEV), vis.to(DEV))
    assert torch.equal(out16, again)                                   # deterministic
    e_ref = (out16[:1].cpu() - ref).abs().max().item()
    e_32 = (out16 - out32).abs().max().item()
    assert e_ref < 1e-2 and e_32 < 1e-2, (e_ref, e_32)
    assert launches16 > 20
    net.conv_engine = 'direct'
    with pytest.raises(RuntimeError), torch.no_grad():
        net(ir.to(DEV), vis.to(DEV))
    net.conv_engine, net.storage = 'auto', 'fp32'
    x = ir.to(DEV).requires_grad_(True)
    net(x, vis.to(DEV)).sum().backward()                               # fp32 mode still differentiates
    assert x.grad is not None


@pytest.mark.parametrize("case", GOLDEN_CASES[:3])
def test_bf16_storage_input_gradients_match_reference_golden(case):
    """storage='bf16' with requires_grad: the forward saves bf16 activations, the backward widens them and runs the
    fp32 (TF32) gradient chain.  Gates of BASELINE.md 6 / SURVEY 8d for the bf16 tier against the reference-autograd
    gradients of the golden fixtures: rel-L2 <= 0.15, sign agreement >= 98 %."""
    from test_gpu_backward import rel_l2, sign_agreement
    g = load_golden(case)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
    net.load_state_dict(g["state_dict"], strict=True)
    net = net.to(DEV).eval()
    net.storage = 'bf16'
    ir = g["ir"].to(DEV).requires_grad_(True)
    vis_full = g["vis"].to(DEV).requires_grad_(True)
    out = net(ir, strided_vis(vis_full))
    assert (out.detach().cpu() - g["out"]).abs().max().item() < 1e-2
    out.backward(g["grad_out"].to(DEV))
    for got, ref in ((ir.grad.cpu(), g["grad_ir"]), (vis_full.grad.cpu()[:, 0:1], g["grad_vis"][:, 0:1])):
        r, s_ = rel_l2(got, ref), sign_agreement(got, ref)
        assert r < 0.15 and s_ > 0.98, (case, r, s_)
