"""CPU: host-side logic of the drop-in — state_dict compatibility, default-init equality with the
reference, weight folding algebra, C-ABI library loads and exports every declared symbol."""
import ctypes
import io
import contextlib
import os
import re

import pytest
import torch
import torch.nn.functional as F

import paif_b200
from paif_b200 import _lib, fusion
from paif_testutil import ROOT, load_golden


def _net(seed=0):
    torch.manual_seed(seed)
    return paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)


def test_state_dict_keys_and_shapes_match_reference():
    ref_sd = load_golden("seed0_default_2x40x56")["state_dict"]
    sd = _net().state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k in sd:
        assert sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype, k


def test_default_init_is_bitwise_the_reference_init():
    # same module tree and construction order => same RNG stream as core/model_fusion_auto.py:600-623
    ref_sd = load_golden("seed0_default_2x40x56")["state_dict"]
    sd = _net(0).state_dict()
    for k in sd:
        assert torch.equal(sd[k], ref_sd[k]), k


def test_reference_checkpoint_loads_strict():
    g = load_golden("seed1_random_1x48x72")
    net = _net(5)
    net.load_state_dict(g["state_dict"], strict=True)


def test_ctor_contract():
    net = _net()
    assert net._C == 32 and net._steps == 4 and net._multiplier == 3 and net._criterion is None
    assert hasattr(net, "_loss") and len(list(net.parameters())) == 39
    with pytest.raises(NotImplementedError):
        paif_b200.Network_Fusion_Searched(16, None, paif_b200.fusion_at)
    bad = paif_b200.fusion_at._replace(normal_3=[('SelAttention_3_1', 0)])     # O((HW)^2) attention: not provided
    with pytest.raises(NotImplementedError):
        paif_b200.Network_Fusion_Searched(32, None, bad)
    bad = paif_b200.fusion_at._replace(normal_3=[('Denseblocks_3_3', 0)])      # (k, d) the reference pads with 0
    with pytest.raises(NotImplementedError):
        paif_b200.Network_Fusion_Searched(32, None, bad)


def test_alternate_genotype_state_dict_matches_reference_fixture():
    from paif_testutil import golden_genotype
    g = load_golden("alt_seed2_random_2x36x52")
    net = paif_b200.Network_Fusion_Searched(32, None, golden_genotype(g))
    ours = net.state_dict()
    assert list(ours.keys()) == list(g["state_dict"].keys())
    for k, v in g["state_dict"].items():
        assert ours[k].shape == v.shape and ours[k].dtype == v.dtype, k
    net.load_state_dict(g["state_dict"], strict=True)


def test_no_cpu_path_and_eval_only():
    net = _net()
    x = torch.rand(1, 1, 16, 16)
    with pytest.raises(RuntimeError):
        net(x, x)                    # training mode
    net.eval()
    with pytest.raises(RuntimeError):
        net(x, x)                    # CPU tensors


def test_decomp_1x1_fold_is_exact():
    torch.manual_seed(0)
    w = torch.randn(32, 128, 1, 1, dtype=torch.float64)
    lf1, lf2, z = (torch.randn(2, 32, 5, 7, dtype=torch.float64) for _ in range(3))
    ref = F.conv2d(torch.cat([lf1, lf2, z - lf1, z - lf2], 1), w)
    got = F.conv2d(torch.cat([lf1, lf2, z], 1), fusion._fold_decomp_1x1(w.float()).double())
    assert (ref - got).abs().max().item() < 1e-5


def test_stem_out_merge_matches_two_padded_convs():
    torch.manual_seed(0)
    w1 = torch.randn(16, 32, 3, 3)
    w2 = torch.randn(1, 16, 3, 3)
    x = torch.randn(1, 32, 9, 11, dtype=torch.float64)
    ref = F.conv2d(F.conv2d(x, w1.double(), None, 1, 1), w2.double(), None, 1, 1)[0, 0]
    wm = fusion._merge_stem_out(w1, w2).double().reshape(3, 3, 5, 5, 32)
    xp = F.pad(x, (2, 2, 2, 2))[0]
    H, W = 9, 11
    got = torch.zeros(H, W, dtype=torch.float64)
    for y in range(H):
        for xx in range(W):
            cy = 0 if y == 0 else (2 if y == H - 1 else 1)
            cx = 0 if xx == 0 else (2 if xx == W - 1 else 1)
            win = xp[:, y:y + 5, xx:xx + 5]                      # [C,5,5]
            got[y, xx] = (win.permute(1, 2, 0) * wm[cy, cx]).sum()
    assert (ref - got).abs().max().item() < 1e-4


def test_dgrad_groups_are_the_conv_transpose():
    torch.manual_seed(0)
    w = torch.randn(32, 64, 3, 3, dtype=torch.float64)
    x = torch.randn(1, 64, 8, 9, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, 1, 2, 2)
    g = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, g)
    groups = fusion._dgrad_groups(w.float(), 3, 2)
    for gi, cw in enumerate(groups):
        # unpack direct layout [1][taps][cin][cout] back to OIHW and run it as a forward conv
        wd = cw.direct[0].double().permute(2, 1, 0).reshape(32, 32, 3, 3)
        got = F.conv2d(g, wd, None, 1, 2, 2)
        assert (got - gx[:, 32 * gi:32 * gi + 32]).abs().max().item() < 1e-4


def test_install_rebinds_reference_symbol():
    import types
    m = types.ModuleType("fake_core_model_fusion_auto")
    m.Network_Fusion_Searched = object
    paif_b200.install(m)
    assert m.Network_Fusion_Searched is paif_b200.Network_Fusion_Searched
    assert not hasattr(m, "Network_Fusion_Searched_showfeatures")
    m.Network_Fusion_Searched_showfeatures = object
    paif_b200.install(m)
    assert m.Network_Fusion_Searched_showfeatures is paif_b200.Network_Fusion_Searched_showfeatures


def test_showfeatures_variant_has_the_reference_state_dict():
    """core/model_fusion_auto.py:641-697 builds the same sub-module tree (Cell_Decom_decom holds what Cell_Decom holds):
    checked in the build container against the reference class itself — keys AND default-init values equal the
    seed-0 fixture of Network_Fusion_Searched."""
    torch.manual_seed(0)
    net = paif_b200.Network_Fusion_Searched_showfeatures(32, None, paif_b200.fusion_at)
    ref_sd = load_golden("seed0_default_2x40x56")["state_dict"]
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys()) and all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    assert hasattr(net, "forward2")


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "paif_b200.h")).read()
    declared = set(re.findall(r"\b(paif_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "library lacks %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.paif_abi_version() == _lib.ABI_VERSION == 5
    assert ctypes.sizeof(_lib.ConvDesc) > 0


def test_weight_images_follow_the_documented_tile_layout():
    """PaifConvDesc.weight_mma as documented in include/paif_b200.h: [K-group][dx][K step][16-B chunk][dy][cout][cin in
    chunk], TF32-rounded fp32 (4 cin per chunk) and bf16 (8 cin per chunk)."""
    torch.manual_seed(0)
    w = torch.randn(32, 64, 3, 3)
    cw = fusion._ConvW(w, 2, 3, 1)
    assert cw.mma.shape == (2, 3, 4, 2, 3, 32, 4) and cw.mma.dtype == torch.float32
    assert cw.mma16.shape == (2, 3, 2, 2, 3, 32, 8) and cw.mma16.dtype == torch.bfloat16
    wt = fusion._round_tf32(w)
    for (g, dx, ks, ch, dy, co, c) in [(0, 0, 0, 0, 0, 0, 0), (1, 2, 3, 1, 1, 17, 2), (0, 1, 2, 0, 2, 31, 3), (1, 0, 1, 1, 0, 5, 1)]:
        assert cw.mma[g, dx, ks, ch, dy, co, c] == wt[co, g * 32 + ks * 8 + ch * 4 + c, dy, dx]
    for (g, dx, ks, ch, dy, co, c) in [(0, 0, 0, 0, 0, 0, 0), (1, 2, 1, 1, 1, 17, 6), (0, 1, 1, 0, 2, 31, 7), (1, 0, 0, 1, 0, 5, 3)]:
        assert cw.mma16[g, dx, ks, ch, dy, co, c] == w[co, g * 32 + ks * 16 + ch * 8 + c, dy, dx].to(torch.bfloat16)
    # TF32 rounding: nearest, 10-bit mantissa kept in an fp32 container
    assert ((wt.view(torch.int32) & 0x1FFF) == 0).all() and (wt - w).abs().max() <= w.abs().max() * 2.0 ** -11


def test_bf16_storage_mode_rules():
    """storage='bf16' is tensor-core only and covers the primitives of the shipped genotype; since round 2 it also
    runs with saved activations (the backward widens them to fp32)."""
    net = _net().eval()
    assert net.storage == 'fp32' and net._bf16_storage(False) is False and net._bf16_storage(True) is False
    net.storage = 'bf16'
    assert net._bf16_storage(False) is True and net._bf16_storage(True) is True
    net.conv_engine = 'direct'
    with pytest.raises(RuntimeError):
        net._bf16_storage(False)
    net.conv_engine, net.storage = 'auto', 'fp16'
    with pytest.raises(ValueError):
        net._bf16_storage(False)
    with contextlib.redirect_stdout(io.StringIO()):
        alt = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at._replace(normal_3=[('SepConv_3_1', 0)]))
    alt.storage = 'bf16'
    with pytest.raises(NotImplementedError):
        alt._bf16_storage(False)


def test_rebuild_is_decided_by_source_content_not_file_times(tmp_path, monkeypatch):
    """paif_b200.build.needs_build: a library that travelled with a snapshot of the tree (file times lost) is used as
    is when the hash written next to it matches the sources, and rebuilt when a source changed."""
    import os
    from paif_b200 import build
    csrc = tmp_path / "csrc"
    csrc.mkdir()
    (csrc / "a.cu").write_text("// kernel a\n")
    inc = tmp_path / "include"
    inc.mkdir()
    (inc / "paif_b200.h").write_text("// header\n")
    pkg = tmp_path / "pkg"
    pkg.mkdir()
    lib = pkg / "libpaif_b200.so"
    monkeypatch.setattr(build, "CSRC", str(csrc))
    monkeypatch.setattr(build, "HERE", str(pkg))
    monkeypatch.setattr(build, "LIB", str(lib))
    assert build.needs_build()                                  # no library yet
    lib.write_bytes(b"\x7fELF")
    (pkg / "libpaif_b200.so.srchash").write_text(build._src_hash() + "\n")
    os.utime(str(csrc / "a.cu"), None)                          # a newer file time alone does not trigger a rebuild
    assert not build.needs_build()
    (csrc / "a.cu").write_text("// kernel a, edited\n")
    assert build.needs_build()


def test_direct_guide_term_of_the_fused_adjoint_is_a_dot_with_mean2_of_the_mixed_A():
    """The algebra behind paif_gf_mix_forward_save / paif_gf_decomp_backward_saved, in fp64 on the oracle's operators:
    with out = conv1x1(cat[LF1, LF2, z - LF1, z - LF2]) (core/model_fusion_auto.py:509-535) and the guided filter
    LF_e = mean2(A_e) g + mean2(b_e), the derivative of out w.r.t. the guide g AT FIXED window statistics is
    mean2(A'_o), A' = Wa A_1 + Wb A_2 (the folded 1x1 applied between the two box-filter levels).  So the adjoint's direct
    term sum_c sum_e gLF_e,c mean2(A_e,c) equals sum_o gx_o mean2(A'_o): one saved map and a dot product."""
    from oracle import fusion_oracle as fo
    torch.manual_seed(11)
    B, C, H, W = 1, 32, 24, 28
    z = torch.rand(B, C, H, W, dtype=torch.float64)
    g = fo.get_residue(z)
    w = torch.randn(C, 4 * C, 1, 1, dtype=torch.float64) * 0.2
    gx = torch.randn(B, C, H, W, dtype=torch.float64)
    wa, wb, _ = fusion._fold_decomp_1x1(w, double=True)
    N = fo.box_filter(g.new_ones((1, 1, H, W)))
    mx = fo.box_filter(g) / N
    var = fo.box_filter(g * g) / N - mx * mx
    mz = fo.box_filter(z) / N
    cov = fo.box_filter(g * z) / N - mx * mz
    mean_a = [fo.box_filter(cov / (var + eps)) / N for eps in (1e-3, 1e-4)]
    # three-pass adjoint: gLF_e = W_e^T gx, direct term = sum_c sum_e gLF_e,c mean2(A_e,c)
    glf = [torch.einsum("oc,bohw->bchw", we, gx) for we in (wa, wb)]
    direct_old = sum((gl * ma).sum(1) for gl, ma in zip(glf, mean_a))
    # fused form: A' = Wa A_1 + Wb A_2 mixed BEFORE the level-2 mean (linearity), direct term = sum_o gx_o mean2(A'_o)
    a_mixed = sum(torch.einsum("oc,bchw->bohw", we, cov / (var + eps)) for we, eps in ((wa, 1e-3), (wb, 1e-4)))
    direct_new = (gx * (fo.box_filter(a_mixed) / N)).sum(1)
    assert torch.allclose(direct_old, direct_new, rtol=1e-10, atol=1e-12)
    # and it IS the partial derivative: perturb g only where it multiplies mean2(A) (autograd with the statistics detached)
    gd = g.clone().requires_grad_(True)
    lf = [ma.detach() * gd + (fo.box_filter(mz - (cov / (var + eps)) * mx) / N).detach() for ma, eps in zip(mean_a, (1e-3, 1e-4))]
    out = torch.nn.functional.conv2d(torch.cat([lf[0], lf[1], z - lf[0], z - lf[1]], 1), w)
    (gg,) = torch.autograd.grad(out, gd, gx)
    assert torch.allclose(gg[:, 0], direct_new, rtol=1e-9, atol=1e-11)
