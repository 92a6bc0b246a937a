"""CPU: the oracle restatement (oracle/fusion_oracle.py) against the golden vectors produced by
the unmodified reference (oracle/gen_golden.py), plus oracle self-consistency checks."""
import torch

from oracle import fusion_oracle as fo
from paif_testutil import golden_genotype


def test_oracle_matches_reference_golden(golden):
    out = fo.fusion_forward(golden["state_dict"], golden_genotype(golden), golden["ir"], golden["vis"])
    assert out.shape == golden["out"].shape
    assert (out - golden["out"]).abs().max().item() <= 1e-6


def test_oracle_input_grads_match_reference_golden(golden):
    _, g_ir, g_vis = fo.fusion_input_grads(golden["state_dict"], golden_genotype(golden), golden["ir"], golden["vis"],
                                           golden["grad_out"])
    for a, b in ((g_ir, golden["grad_ir"]), (g_vis, golden["grad_vis"])):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item())
    # only the Y channel of vis receives gradient (core/model_fusion_auto.py:626)
    assert golden["grad_vis"][:, 1:].abs().max().item() == 0.0


def test_state_dict_has_the_45_reference_keys(golden):
    if "genotype" in golden:
        return                      # alternate-genotype fixture: different key set
    sd = golden["state_dict"]
    assert len(sd) == 45
    assert sum(v.numel() for v in sd.values()) == 260002
    assert sd["decompation.conv1x1_lf.weight"].shape == (32, 128, 1, 1)
    assert sd["chain._ops.1._op.op.0.conv.weight"].shape == (32, 32, 7, 7)


def test_box_filter_is_clipped_window_sum():
    x = torch.rand(1, 2, 23, 31, dtype=torch.float64)
    ref = torch.nn.functional.avg_pool2d(x, 9, 1, 4, count_include_pad=True) * 81
    assert (fo.box_filter(x, 4) - ref).abs().max().item() < 1e-10


def test_confusion_matrix_matches_sklearn():
    from sklearn.metrics import confusion_matrix
    g = torch.Generator().manual_seed(0)
    label = torch.randint(0, 10, (5000,), generator=g)       # includes an out-of-range class 9
    pred = torch.randint(0, 9, (5000,), generator=g)
    ours = fo.confusion_matrix(label, pred, 9).numpy()
    ref = confusion_matrix(label.numpy(), pred.numpy(), labels=list(range(9)))
    assert (ours == ref).all()
