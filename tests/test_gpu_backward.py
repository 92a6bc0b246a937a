"""GPU: backward-to-input (what attack/attack.py:501's loss.backward() needs from the fusion net)
against the reference's autograd gradients stored in the golden fixtures and against autograd of the
CPU oracle.  The gradient is ill-conditioned (arg-max/min guide, PReLU kinks, 1/(var+eps)), so the gate
is rel-L2 + sign agreement, not max-abs."""
import ctypes

import pytest
import torch

import paif_b200
from oracle import fusion_oracle as fo
from paif_b200 import _lib
from paif_testutil import GOLDEN_CASES, golden_genotype, load_golden, strided_vis

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


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def sign_agreement(a, b):
    m = b.abs() > 1e-3 * b.abs().max()
    return (torch.sign(a[m]) == torch.sign(b[m])).float().mean().item()


@pytest.mark.parametrize("shape,smooth", [((1, 40, 56), False), ((2, 33, 47), True)])
def test_guided_filter_adjoint(shape, smooth):
    B, H, W = shape
    torch.manual_seed(3)
    z = torch.rand(B, 32, H, W, dtype=torch.float64)
    if smooth:
        z = torch.nn.functional.avg_pool2d(z, 9, 1, 4)
    guide = fo.get_residue(z).detach().requires_grad_(True)
    zr = z.clone().requires_grad_(True)
    lf1 = fo.guided_filter(guide, zr, 4, 1e-3)
    lf2 = fo.guided_filter(guide, zr, 4, 1e-4)
    g1, g2 = torch.randn_like(lf1), torch.randn_like(lf2)
    gz, gg = torch.autograd.grad([lf1, lf2], [zr, guide], [g1, g2])
    zc, gc = to_c4(z.float()).to(DEV), guide.detach().float()[:, 0].contiguous().to(DEV)
    g1c, g2c = to_c4(g1.float()).to(DEV), to_c4(g2.float()).to(DEV)
    gfeat = torch.empty_like(zc)
    gres = torch.empty(_lib.load().paif_gf_guide_parts(32), B, H, W, device=DEV)
    stats = torch.empty(3, B, H, W, device=DEV)
    _lib.call("paif_gf_guide_stats", gc.data_ptr(), stats.data_ptr(), B, H, W, stream())
    work = torch.empty(_lib.load().paif_gf_backward_work_floats(32, B, H, W), device=DEV)
    _lib.call("paif_gf_decomp_backward", zc.data_ptr(), gc.data_ptr(), stats.data_ptr(), g1c.data_ptr(), g2c.data_ptr(),
              gfeat.data_ptr(), gres.data_ptr(), work.data_ptr(), 32, B, H, W, stream())
    e_z = rel_l2(from_c4(gfeat).cpu().double(), gz)
    e_g = rel_l2(gres.sum(0).cpu().double(), gg[:, 0])
    assert e_z < 2e-3 and e_g < 2e-3, (e_z, e_g)


@pytest.mark.parametrize("shape,smooth", [((1, 40, 56), False), ((2, 33, 48), True), ((1, 200, 236), False)])
def test_fused_decomposition_adjoint_from_saved_mean(shape, smooth):
    """paif_gf_mix_forward_save + paif_gf_decomp_backward_saved (the direct guide term from the forward's mean2(A')) against
    fp64 autograd of conv1x1(cat[LF, HF]) (core/model_fusion_auto.py:509-535) and against the forward-recompute adjoint."""
    from paif_b200 import fusion
    B, H, W = shape
    torch.manual_seed(5)
    z = torch.rand(B, 32, H, W, dtype=torch.float64)
    if smooth:
        z = torch.nn.functional.avg_pool2d(z, 9, 1, 4)
    w = torch.randn(32, 128, 1, 1, dtype=torch.float64) * 0.15
    bias = torch.randn(32, dtype=torch.float64) * 0.1
    guide = fo.get_residue(z).detach().requires_grad_(True)
    zr = z.clone().requires_grad_(True)
    lf1, lf2 = fo.guided_filter(guide, zr, 4, 1e-3), fo.guided_filter(guide, zr, 4, 1e-4)
    x = torch.nn.functional.conv2d(torch.cat([lf1, lf2, zr - lf1, zr - lf2], 1), w, bias)
    gx = torch.randn_like(x)
    gz_ref, gg_ref = torch.autograd.grad(x, [zr, guide], gx)
    wa, wb, wc = fusion._fold_decomp_1x1(w, double=True)
    g1 = torch.einsum("oc,bohw->bchw", wa, gx)          # the transposed 1x1 (three K groups), as the module's backward forms them
    g2 = torch.einsum("oc,bohw->bchw", wb, gx)
    gzd = torch.einsum("oc,bohw->bchw", wc, gx)
    zc, gc = to_c4(z.float()).to(DEV), guide.detach().float()[:, 0].contiguous().to(DEV)
    g1c, g2c, gxc = to_c4(g1.float()).to(DEV), to_c4(g2.float()).to(DEV), to_c4(gx.float()).to(DEV)
    stats = torch.empty(3, B, H, W, device=DEV)
    _lib.call("paif_gf_guide_stats", gc.data_ptr(), stats.data_ptr(), B, H, W, stream())
    wp, bd = fusion._pack_gf_mix(w.float().to(DEV)), bias.float().to(DEV)
    out, out2, ma = torch.empty_like(zc), torch.empty_like(zc), torch.empty_like(zc)
    _lib.call("paif_gf_mix_forward_save", zc.data_ptr(), gc.data_ptr(), stats.data_ptr(), wp.data_ptr(), bd.data_ptr(),
              out.data_ptr(), 0, ma.data_ptr(), 32, B, H, W, stream())
    _lib.call("paif_gf_mix_forward", zc.data_ptr(), gc.data_ptr(), stats.data_ptr(), wp.data_ptr(), bd.data_ptr(),
              out2.data_ptr(), 0, 32, B, H, W, stream())
    assert torch.equal(out, out2)                        # saving mean2(A') does not change the output
    nparts = _lib.load().paif_gf_guide_parts(32)
    work = torch.empty(_lib.load().paif_gf_backward_work_floats(32, B, H, W), device=DEV)
    res = []
    for saved in (True, False):
        gfeat, gres = torch.empty_like(zc), torch.empty(nparts, B, H, W, device=DEV)
        if saved:
            _lib.call("paif_gf_decomp_backward_saved", zc.data_ptr(), gc.data_ptr(), stats.data_ptr(), g1c.data_ptr(), g2c.data_ptr(),
                      gxc.data_ptr(), ma.data_ptr(), gfeat.data_ptr(), gres.data_ptr(), work.data_ptr(), 32, B, H, W, stream())
        else:
            _lib.call("paif_gf_decomp_backward", zc.data_ptr(), gc.data_ptr(), stats.data_ptr(), g1c.data_ptr(), g2c.data_ptr(),
                      gfeat.data_ptr(), gres.data_ptr(), work.data_ptr(), 32, B, H, W, stream())
        res.append((from_c4(gfeat).cpu().double() + gzd, gres.sum(0).cpu().double()))
    (gz_s, gg_s), (gz_r, gg_r) = res
    assert torch.equal(gz_s, gz_r)                       # the feature gradient does not involve the direct term
    e_z, e_g, e_gr = rel_l2(gz_s, gz_ref), rel_l2(gg_s, gg_ref[:, 0]), rel_l2(gg_r, gg_ref[:, 0])
    # TF32 operand rounding of the mix inside mean2(A') (the forward's own rounding) bounds the guide-gradient error
    assert e_z < 2e-3 and e_g < 5e-3 and e_gr < 2e-3, (e_z, e_g, e_gr)


# (engine, fused-image gate, gradient rel-L2 gate, sign-agreement gate): the exact-fp32 engine is held to the
# reference's own fp32-vs-fp64 noise floor, the TF32 tensor-core engine to SURVEY.md 8d's TF32 gates
GRAD_GATES = [("direct", 5e-5, 1e-2, 0.999), ("tcgen05", 1e-3, 5e-2, 0.995)]


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("engine,out_tol,l2_tol,sign_tol", GRAD_GATES)
def test_input_gradients_match_reference_golden(case, engine, out_tol, l2_tol, sign_tol):
    g = load_golden(case)
    net = paif_b200.Network_Fusion_Searched(32, None, golden_genotype(g))
    net.load_state_dict(g["state_dict"], strict=True)
    net = net.to(DEV).eval()
    net.conv_engine = engine
    ir = g["ir"].to(DEV).requires_grad_(True)
    vis_full = g["vis"].to(DEV).requires_grad_(True)
    out = net(ir, strided_vis(vis_full))
    assert (out.detach().cpu() - g["out"]).abs().max().item() < out_tol
    out.backward(g["grad_out"].to(DEV))
    g_ir, g_vis = ir.grad.cpu(), vis_full.grad.cpu()
    assert g_vis[:, 1:].abs().max().item() == 0.0          # only Y gets gradient
    for got, ref in ((g_ir, g["grad_ir"]), (g_vis[:, 0:1], g["grad_vis"][:, 0:1])):
        r, s = rel_l2(got, ref), sign_agreement(got, ref)
        assert r < l2_tol and s > sign_tol, (case, engine, r, s)


def test_no_grad_forward_saves_nothing_and_grad_only_where_needed():
    g = load_golden("seed0_default_2x40x56")
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
    net.load_state_dict(g["state_dict"])
    net = net.to(DEV).eval()
    ir = g["ir"].to(DEV).requires_grad_(True)
    vis = g["vis"].to(DEV)                                  # no grad wanted for vis
    out = net(ir, vis)
    out.sum().backward()
    assert ir.grad is not None and vis.grad is None
    with torch.no_grad():
        out2 = net(ir, vis)
    assert not out2.requires_grad and torch.equal(out2, out.detach())
    assert all(p.grad is None for p in net.parameters())    # weight gradients are intentionally not produced


@pytest.mark.parametrize("shape", [(1, 24, 40), (2, 120, 200)])
def test_forward_backward_bit_deterministic(shape):
    """No atomics, fixed-order partial sums: repeated forward+backward calls return identical bits (both DilConv
    lowerings: dense on the tensor-core engine, and the FFMA depthwise+1x1 kernel)."""
    B, H, W = shape
    torch.manual_seed(0)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    vis, ir = torch.rand(B, 3, H, W, generator=g).to(DEV), torch.rand(B, 1, H, W, generator=g).to(DEV)
    cot = torch.randn(B, 1, H, W, generator=g).to(DEV)
    outs = {}
    for dense in (False, True):
        net.dilconv_dense = dense
        runs = []
        for _ in range(3):
            a, v = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
            out = net(a, v)
            out.backward(cot)
            runs.append((out.detach(), a.grad, v.grad))
        for r in runs[1:]:
            assert all(torch.equal(x, y) for x, y in zip(runs[0], r))
        outs[dense] = runs[0]
    # the two lowerings agree to TF32-operand accuracy
    assert (outs[True][0] - outs[False][0]).abs().max().item() < 1e-3


def test_full_size_forward_and_gradients_480x640():
    """BASELINE shape (480x640): forward against the CPU oracle, input gradients against the oracle's autograd,
    default ('auto' = tcgen05 TF32) engine; plus batch-position invariance with the full-size tiling (batch 17
    makes the conv engine use 32-row chunks and TMEM slot reuse)."""
    from oracle import fusion_oracle as fo
    g = load_golden("seed1_random_1x48x72")
    sd = g["state_dict"]
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    gen = torch.Generator().manual_seed(21)
    ir, vis = torch.rand(1, 1, 480, 640, generator=gen), torch.rand(1, 3, 480, 640, generator=gen)
    cot = torch.randn(1, 1, 480, 640, generator=gen)
    ir_r, vis_r = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
    ref = fo.fusion_forward(sd, paif_b200.fusion_at, ir_r, vis_r)
    g_ir_ref, g_vis_ref = torch.autograd.grad(ref, [ir_r, vis_r], cot)
    ir_d, vis_d = ir.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    out = net(ir_d, vis_d)
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 1e-3
    out.backward(cot.to(DEV))
    for got, want in ((ir_d.grad.cpu(), g_ir_ref), (vis_d.grad.cpu()[:, 0:1], g_vis_ref[:, 0:1])):
        r, s = rel_l2(got, want), sign_agreement(got, want)
        assert r < 5e-2 and s > 0.995, (r, s)
    with torch.no_grad():
        big = net(ir.to(DEV).expand(17, -1, -1, -1).contiguous(), vis.to(DEV).expand(17, -1, -1, -1).contiguous())
    # (the guided-filter row chunking adapts to the batch size, so batch 17 vs 1 is close, not bit-identical)
    assert (big[0] - out.detach()[0]).abs().max().item() <= 1e-4
    assert torch.equal(big[0], big[16])
