"""GPU: end-to-end parity of the drop-in module against (a) the golden vectors produced by the
unmodified reference and (b) the CPU oracle on fresh seeded inputs.  Tolerances: north_star's
max-abs 1e-3 on the fused image in fp32 mode; the exact-fp32 direct engine is held to 5e-5."""
import pytest
import torch

import paif_b200
from oracle import fusion_oracle as fo
from paif_testutil import GOLDEN_CASES, golden_genotype, load_golden, strided_vis

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ENGINES = [("direct", 5e-5), ("tcgen05", 1e-3)]      # (conv engine, max-abs gate on the fused image)


def build(sd, engine, genotype=paif_b200.fusion_at):
    net = paif_b200.Network_Fusion_Searched(32, None, genotype)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    net.conv_engine = engine
    return net


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("engine,tol", ENGINES)
def test_forward_matches_reference_golden(case, engine, tol):
    g = load_golden(case)
    net = build(g["state_dict"], engine, golden_genotype(g))
    with torch.no_grad():
        out = net(g["ir"].to(DEV), strided_vis(g["vis"].to(DEV)))
    assert out.shape == g["out"].shape and out.is_contiguous()
    err = (out.cpu() - g["out"]).abs().max().item()
    assert err <= tol, err
    assert net.last_launches > 0


@pytest.mark.parametrize("engine,tol", ENGINES)
def test_forward_matches_oracle_on_fresh_inputs(engine, tol):
    g = load_golden("seed1_random_1x48x72")
    net = build(g["state_dict"], engine)
    torch.manual_seed(11)
    ir, vis = torch.rand(2, 1, 64, 136), torch.rand(2, 3, 64, 136)
    ref = fo.fusion_forward(g["state_dict"], paif_b200.fusion_at, ir, vis)
    with torch.no_grad():
        out = net(ir.to(DEV), vis.to(DEV))
    assert (out.cpu() - ref).abs().max().item() <= tol


def test_batch_position_invariance_and_determinism():
    g = load_golden("seed0_default_2x40x56")
    net = build(g["state_dict"], "direct")
    ir, vis = g["ir"].to(DEV), g["vis"].to(DEV)
    with torch.no_grad():
        both = net(ir, vis)
        again = net(ir, vis)
        one = net(ir[1:2], vis[1:2])
    assert torch.equal(both, again)
    assert torch.equal(both[1:2], one)


@pytest.mark.parametrize("shape", [(1, 10, 10), (3, 11, 300), (1, 130, 10), (2, 77, 203)])
def test_edge_sizes_forward_and_input_gradients(shape):
    """Smallest legal image (the guided filter needs H, W > 9), wide-and-short, tall-and-narrow, and a size that is
    a multiple of nothing: forward within 1e-3 and input gradients within the TF32 gates of the CPU oracle."""
    B, H, W = shape
    g = load_golden("seed1_random_1x48x72")
    net = build(g["state_dict"], "auto")
    gen = torch.Generator().manual_seed(31)
    ir, vis = torch.rand(B, 1, H, W, generator=gen), torch.rand(B, 3, H, W, generator=gen)
    cot = torch.randn(B, 1, H, W, generator=gen)
    ref, g_ir, g_vis = fo.fusion_input_grads(g["state_dict"], paif_b200.fusion_at, ir, vis, cot)
    a, v = ir.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    out = net(a, v)
    assert (out.detach().cpu() - ref).abs().max().item() <= 1e-3
    out.backward(cot.to(DEV))
    for got, want in ((a.grad.cpu(), g_ir), (v.grad.cpu()[:, 0:1], g_vis[:, 0:1])):
        assert ((got - want).norm() / want.norm()).item() < 1e-1      # gross-error gate; the calibrated TF32 gates are in test_gpu_backward.py


def test_forward2_returns_the_reference_intermediates():
    """Network_Fusion_Searched_showfeatures.forward2 (core/model_fusion_auto.py:669-679) against the oracle's
    intermediates (which equal the reference's forward2 outputs bit for bit on CPU, checked in the build container)."""
    g = load_golden("seed1_random_1x48x72")
    sd = g["state_dict"]
    net = paif_b200.Network_Fusion_Searched_showfeatures(32, None, paif_b200.fusion_at)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    net.conv_engine = 'direct'
    ir, vis = g["ir"], g["vis"]
    outs = [t.cpu() for t in net.forward2(ir.to(DEV), vis.to(DEV))]
    inter = {}
    out = fo.fusion_forward(sd, paif_b200.fusion_at, ir, vis, inter)
    want = [out, inter["ir_feature"], inter["vis_feature"]]
    for f in (inter["fir"], inter["fvis"]):
        lf, hf = fo.decomposition(f)
        want += [lf, hf, fo.get_residue(f)]
    assert len(outs) == 9
    for i, (a, b) in enumerate(zip(outs, want)):
        assert a.shape == b.shape, i
        assert (a - b).abs().max().item() < (5e-4 if i in (3, 4, 6, 7) else 5e-5), (i, (a - b).abs().max().item())



@pytest.mark.parametrize("storage,tol", [("fp32", 1e-3), ("bf16", 1e-2)])
def test_config_d_shape_768x1024_matches_oracle(storage, tol):
    """BASELINE configs[3] (M3FD shape 768x1024): the fused image of one pair — computed inside a batch of 3, so the
    multi-chunk / multi-wave tilings of the big shape are exercised — against the CPU oracle, in both storage modes
    (north_star: 1e-3 fp32 / 1e-2 bf16 on the [0,1] scale)."""
    g = load_golden("seed1_random_1x48x72")
    net = build(g["state_dict"], "auto")
    net.storage = storage
    torch.manual_seed(21)
    ir, vis = torch.rand(3, 1, 768, 1024), torch.rand(3, 3, 768, 1024)
    ref = fo.fusion_forward(g["state_dict"], paif_b200.fusion_at, ir[1:2], vis[1:2])
    with torch.no_grad():
        out = net(ir.to(DEV), vis.to(DEV))
    err = (out[1:2].cpu() - ref).abs().max().item()
    assert err <= tol, err


def test_bench_tiling_batch16_480x640_matches_oracle():
    """The launch geometry bench.py times (batch 16 x 480x640: 7 row chunks per conv strip, 3 chunks per guided-filter
    strip, persistent CTAs walking several work items): two pairs of the batch against the CPU oracle."""
    g = load_golden("seed0_default_2x40x56")
    net = build(g["state_dict"], "auto")
    torch.manual_seed(1)
    ir, vis = torch.rand(16, 1, 480, 640), torch.rand(16, 3, 480, 640)
    with torch.no_grad():
        out = net(ir.to(DEV), vis.to(DEV)).cpu()
    for b in (0, 11):
        ref = fo.fusion_forward(g["state_dict"], paif_b200.fusion_at, ir[b:b + 1], vis[b:b + 1])
        err = (out[b:b + 1] - ref).abs().max().item()
        assert err <= 1e-3, (b, err)


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 40, 56), (1, 96, 200)])
def test_whole_network_entry_point_is_bit_identical_to_the_per_operator_path(storage, shape):
    """paif_fusion_forward (one C-ABI call, caller-owned workspace) launches the same kernels in the same order as the
    per-operator path driven from Python: same bits.  Also called here straight through ctypes — device pointers, sizes
    and a stream, no fusion._Runtime — on the strided Y view the reference wrappers pass."""
    import ctypes
    from paif_b200 import _lib
    B, H, W = shape
    g = load_golden("seed1_random_1x48x72")
    net = build(g["state_dict"], "auto")
    net.storage = storage
    torch.manual_seed(5)
    ir, vis = torch.rand(B, 1, H, W).to(DEV), strided_vis(torch.rand(B, 3, H, W).to(DEV))
    with torch.no_grad():
        net.native_forward = False
        per_op = net(ir, vis)
        net.native_forward = True
        native = net(ir, vis)
    assert torch.equal(per_op, native)
    lib = _lib.load()
    st = _lib.STORAGE_BF16 if storage == "bf16" else _lib.STORAGE_F32
    w = net._native_weights(net._packed(False))
    need = lib.paif_fusion_workspace_bytes(B, H, W, st)
    assert need > 0
    ws = torch.empty(need, device=DEV, dtype=torch.uint8)
    out = torch.empty(B, 1, H, W, device=DEV)
    y = vis[:, 0:1]
    rc = lib.paif_fusion_forward(ctypes.byref(w), ir.data_ptr(), ir.stride(0), ir.stride(2), ir.stride(3),
                                 y.data_ptr(), y.stride(0), y.stride(2), y.stride(3), out.data_ptr(), ws.data_ptr(), need,
                                 st, B, H, W, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.paif_last_error_string()
    torch.cuda.synchronize()
    assert torch.equal(out, per_op)
    rc = lib.paif_fusion_forward(ctypes.byref(w), ir.data_ptr(), ir.stride(0), ir.stride(2), ir.stride(3),
                                 y.data_ptr(), y.stride(0), y.stride(2), y.stride(3), out.data_ptr(), ws.data_ptr(), need - 1,
                                 st, B, H, W, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc != 0 and b"workspace" in lib.paif_last_error_string()


@pytest.mark.parametrize("shape", [(2, 40, 56), (1, 96, 200), (3, 33, 132)])
def test_whole_network_backward_entry_points_are_bit_identical_to_the_per_operator_path(shape):
    """paif_fusion_forward_save + paif_fusion_backward_input (the PGD inner step as two C-ABI calls over one caller-owned
    workspace) against the per-operator autograd node: same fused image and same input gradients, bit for bit.  Also
    driven here straight through ctypes — pointers, sizes and a stream, no fusion._Runtime, no autograd."""
    import ctypes
    from paif_b200 import _lib
    B, H, W = shape
    g = load_golden("seed1_random_1x48x72")
    net = build(g["state_dict"], "auto")
    torch.manual_seed(6)
    ir0, vis0 = torch.rand(B, 1, H, W).to(DEV), strided_vis(torch.rand(B, 3, H, W).to(DEV))
    gout = (torch.rand(B, 1, H, W) - 0.5).to(DEV)
    res = {}
    for native in (False, True):
        net.native_forward = native
        a = ir0.clone().requires_grad_(True)
        v = vis0.detach().requires_grad_(True)
        out = net(a, v)
        out.backward(gout)
        res[native] = (out.detach().clone(), a.grad.clone(), v.grad.clone())
    for x, y in zip(res[False], res[True]):
        assert torch.equal(x, y)
    assert res[True][1].abs().max().item() > 0
    # plain ctypes
    lib = _lib.load()
    p = net._packed(True)
    w, gw = net._native_weights(p), net._native_grad_weights(p)
    need = lib.paif_fusion_train_workspace_bytes(B, H, W)
    assert need > 0
    ws = torch.empty(need, device=DEV, dtype=torch.uint8)
    out = torch.empty(B, 1, H, W, device=DEV)
    y = vis0[:, 0:1]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.paif_fusion_forward_save(ctypes.byref(w), ir0.data_ptr(), ir0.stride(0), ir0.stride(2), ir0.stride(3),
                                      y.data_ptr(), y.stride(0), y.stride(2), y.stride(3), out.data_ptr(), ws.data_ptr(),
                                      need, B, H, W, st)
    assert rc == 0, lib.paif_last_error_string()
    g_ir, g_vis = torch.empty(B, H, W, device=DEV), torch.empty(B, H, W, device=DEV)
    out.zero_()                                   # the backward must not depend on the caller's copy of the output
    rc = lib.paif_fusion_backward_input(ctypes.byref(w), ctypes.byref(gw), gout.data_ptr(), g_ir.data_ptr(),
                                        g_vis.data_ptr(), ws.data_ptr(), need, B, H, W, st)
    assert rc == 0, lib.paif_last_error_string()
    torch.cuda.synchronize()
    assert torch.equal(g_ir, res[False][1][:, 0])
    assert torch.equal(g_vis, res[False][2][:, 0])
    rc = lib.paif_fusion_backward_input(ctypes.byref(w), ctypes.byref(gw), gout.data_ptr(), g_ir.data_ptr(),
                                        g_vis.data_ptr(), ws.data_ptr(), need - 1, B, H, W, st)
    assert rc != 0 and b"workspace" in lib.paif_last_error_string()
