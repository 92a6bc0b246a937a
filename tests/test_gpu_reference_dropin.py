"""The drop-in claim, checked with the reference's OWN code around the paif_b200 fusion net:
``Network_MM_Searched`` / ``Network_MM_CompModel`` (core/model_fusion_auto.py:698-729, 1029-1060), ``attack_both``
(attack/attack.py:417-514) and the entry scripts ``robust_test.py`` / ``test_original.py``, all unmodified, once with
the reference fusion class and once with ``paif_b200.install()`` — same weights (``load_state_dict(strict=True)``),
same inputs, same RNG seed.  Needs the reference tree (``/root/reference`` or the staged ``baseline/_ref``)."""
import os

import numpy as np
import pytest
import torch

import reference_harness as rh

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rh.available(), reason="reference tree not present (baseline/_ref not staged)")]


@pytest.fixture(autouse=True)
def _exact_fp32_stock_ops():
    """The stock-PyTorch parts (SegFormer, colour glue; and the whole reference fusion net) run in true fp32 so that
    the only difference between the two arms is the paif_b200 fusion path."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _inputs(B, H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    ir = torch.rand(B, 1, H, W, generator=g).cuda()
    vis = torch.rand(B, 3, H, W, generator=g).cuda()
    label = torch.randint(0, 9, (B, H, W), generator=g).cuda()
    return ir, vis, label


def _pair(kind, seed=0):
    ref = rh.build_reference_task(kind, seed)
    ours = rh.build_dropin_task(kind, ref.state_dict())
    return ref.cuda().eval(), ours.cuda().eval()


@pytest.mark.parametrize("kind", ["searched", "comp"])
@pytest.mark.parametrize("engine,tol,seg_tol", [("direct", 5e-5, 2e-2), ("auto", 1e-3, 1e-1)])
def test_reference_wrappers_around_the_dropin(kind, engine, tol, seg_tol):
    """634 keys load strict; fused image within north_star's fp32 tolerance; the stock SegFormer sees the same image."""
    ref, ours = _pair(kind)
    assert len(ours.state_dict()) == len(ref.state_dict()) == 634
    ours.enhance_net.conv_engine = engine
    for B, H, W in ((1, 96, 128), (2, 64, 160)):
        ir, vis, _ = _inputs(B, H, W)
        with torch.no_grad():
            f_ref, s_ref = ref(ir, vis)
            f_our, s_our = ours(ir, vis)
        assert f_our.shape == f_ref.shape and s_our.shape == s_ref.shape
        assert (f_our - f_ref).abs().max().item() <= tol
        assert (s_our - s_ref).abs().max().item() <= seg_tol * max(1.0, s_ref.abs().max().item())
        assert (s_our.argmax(1) == s_ref.argmax(1)).float().mean().item() >= 0.99
        if kind == "comp":
            with torch.no_grad():
                assert (ours.forward_fusion(ir, vis) - ref.forward_fusion(ir, vis)).abs().max().item() <= tol


@pytest.mark.parametrize("engine,agree", [("direct", 0.995), ("auto", 0.98)])
def test_reference_attack_both_unchanged(engine, agree):
    """``attack.attack.attack_both`` (unmodified) drives the drop-in through autograd exactly as it drives the
    reference model: same seeded start, one PGD step -> the sign steps agree on >= 99.5 % (exact-fp32 engine) /
    98 % (TF32 tensor-core engine) of the pixels; the returned order is (delta_ir, delta_vis)."""
    _, atk = rh.reference_modules()
    ref, ours = _pair("searched", seed=1)
    ours.enhance_net.conv_engine = engine
    ir, vis, label = _inputs(1, 96, 128, seed=3)
    res = []
    for model in (ref, ours):
        torch.manual_seed(11)
        d_ir, d_vis = atk.attack_both(model, X_vis=vis, X_ir=ir, label=label, attack_loss='l_seg', attack_iters=1,
                                      epsilon=8 / 255., alpha=2 / 255.)
        assert d_ir.shape == ir.shape and d_vis.shape == vis.shape
        res.append((d_ir.detach(), d_vis.detach()))
    # only channel 0 of vis reaches the fusion net; channels 1-2 get gradient through the colour glue only
    for got, want in zip(res[1], res[0]):
        same = ((got - want).abs() <= 1e-6).float().mean().item()
        assert same >= agree, same
        assert (got - want).abs().max().item() <= 2 * 2 / 255. + 1e-6
    # a longer attack stays inside the epsilon ball and the [0, 1] box
    torch.manual_seed(12)
    d_ir, d_vis = atk.attack_both(ours, X_vis=vis, X_ir=ir, label=label, attack_loss='l_seg', attack_iters=3)
    assert d_ir.abs().max().item() <= 8 / 255. + 1e-6 and d_vis.abs().max().item() <= 8 / 255. + 1e-6
    assert (ir + d_ir).min().item() >= -1e-6 and (ir + d_ir).max().item() <= 1 + 1e-6


def _count_dropin_forwards(monkeypatch):
    import paif_b200
    calls = {"n": 0}
    orig = paif_b200.Network_Fusion_Searched.forward

    def counted(self, ir, vis):
        calls["n"] += 1
        return orig(self, ir, vis)

    monkeypatch.setattr(paif_b200.Network_Fusion_Searched, "forward", counted)
    return calls


def test_test_original_script_runs_unchanged(tmp_path, monkeypatch):
    """test_original.py (clean evaluation: fusion + SegFormer, PNG outputs, sklearn confusion matrix, result file)
    via ``runpy`` with the drop-in installed; outputs compared with the same script run on the reference class."""
    n = 2
    wd = rh.prepare_workdir(str(tmp_path / "wd"), "test_original.py", n_frames=n, H=96, W=128)
    argv = ["--num_workers", "0"]
    rh.run_script("test_original.py", wd, argv, use_dropin=False)
    root = os.path.join(wd, "attack", "our_orignal_l_seg_PGD5_8_2_both")
    ref_fused, ref_seg = rh.read_pngs(os.path.join(root, "fused_attacked")), rh.read_pngs(os.path.join(root, "seg_attacked"))
    ref_txt = open(os.path.join(root, "our_orignal_PGD5_8_2.txt")).read()
    os.rename(root, root + "_reference")
    calls = _count_dropin_forwards(monkeypatch)
    out = rh.run_script("test_original.py", wd, argv, use_dropin=True)
    assert calls["n"] == n and "model load done" in out
    fused, seg = rh.read_pngs(os.path.join(root, "fused_attacked")), rh.read_pngs(os.path.join(root, "seg_attacked"))
    assert sorted(fused) == sorted(ref_fused) and len(fused) == n
    for name in fused:
        d = np.abs(fused[name].astype(np.int32) - ref_fused[name].astype(np.int32))
        # 1e-3 on [0,1] is a quarter of a grey level; the script then min-max stretches the uint8 image
        # (test_original.py:196-200), which multiplies a one-level difference by 255 / (max - min)
        assert d.max() <= 8 and d.mean() <= 0.25 and (d > 0).mean() <= 0.05
        assert (seg[name] == ref_seg[name]).mean() >= 0.99
    txt = open(os.path.join(root, "our_orignal_PGD5_8_2.txt")).read()
    assert txt.splitlines()[:4] == ref_txt.splitlines()[:4]


def test_robust_test_script_runs_unchanged(tmp_path, monkeypatch):
    """robust_test.py (PGD through fusion + SegFormer via attack_both, then the attacked forward) with the drop-in
    installed.  PGD walks along sign(grad), so the two runs are compared statistically."""
    n, iters = 2, 2
    wd = rh.prepare_workdir(str(tmp_path / "wd"), "robust_test.py", n_frames=n, H=96, W=128)
    argv = ["--num_workers", "0", "--attack_iters", str(iters)]
    rh.run_script("robust_test.py", wd, argv, use_dropin=False, seed=5)
    root = os.path.join(wd, "attack", "meta_final_l_seg_PGD%d_8_2_both" % iters)
    ref_ir, ref_fused = rh.read_pngs(os.path.join(root, "ir_attacked")), rh.read_pngs(os.path.join(root, "fused_attacked"))
    os.rename(root, root + "_reference")
    calls = _count_dropin_forwards(monkeypatch)
    rh.run_script("robust_test.py", wd, argv, use_dropin=True, seed=5)
    assert calls["n"] == n * (iters + 1)                    # `iters` attack forwards + the attacked forward per frame
    ir, fused = rh.read_pngs(os.path.join(root, "ir_attacked")), rh.read_pngs(os.path.join(root, "fused_attacked"))
    assert os.path.isfile(os.path.join(root, "meta_final_PGD%d_8_2.txt" % iters))
    for name in ir:
        assert (ir[name] == ref_ir[name]).mean() >= 0.9                      # same start, mostly the same sign steps
        d = np.abs(fused[name].astype(np.int32) - ref_fused[name].astype(np.int32))
        assert d.mean() <= 2.0
