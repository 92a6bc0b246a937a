"""GPU: the wrapper glue either side of the fusion path as kernels — RGB -> Y inside the visible
stem (``forward_rgb``) and the output-side colour / normalisation glue (``paif_glue_forward`` / ``_backward``) — against
the stock-PyTorch restatement of core/model_fusion_auto.py:69-111, 712-728 (``FusionSegTask(fused_glue=False)``, itself
checked against the reference formulas in tests/test_eval_host.py) and, when the reference tree is staged, against the
reference wrapper ``Network_MM_CompModel`` itself."""
import pytest
import torch
import torch.nn as nn

import paif_b200
from paif_b200.consumer import FusionSegTask, _GlueFn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class ProbeConsumer(nn.Module):
    """returns its input: the normalised image the consumer would see"""

    def forward(self, x):
        return x


def _nets(per_sample):
    torch.manual_seed(0)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
    net.conv_engine = 'direct'
    a = FusionSegTask(net, ProbeConsumer(), per_sample_minmax=per_sample, fused_glue=False).to(DEV).eval()
    b = FusionSegTask(net, ProbeConsumer(), per_sample_minmax=per_sample, fused_glue=True).to(DEV).eval()
    return a, b


@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("shape", [(1, 40, 56), (3, 33, 47)])
def test_fused_glue_forward_and_gradients(per_sample, shape):
    B, H, W = shape
    stock, fusedk = _nets(per_sample)
    g = torch.Generator().manual_seed(3)
    ir = torch.rand(B, 1, H, W, generator=g).to(DEV)
    vis = (torch.rand(B, 3, H, W, generator=g) * 1.4 - 0.2).to(DEV)           # some pixels clamp at 0 and at 1
    cot = torch.randn(B, 3, H, W, generator=g).to(DEV)
    res = []
    for task in (stock, fusedk):
        a, v = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
        fused, x = task(a, v)
        x.backward(cot)
        res.append((fused.detach(), x.detach(), a.grad, v.grad))
    (f0, x0, ga0, gv0), (f1, x1, ga1, gv1) = res
    assert (f1 - f0).abs().max().item() <= 1e-6                 # RGB -> Y in the stem: the same Y to the last bit or two
    assert (x1 - x0).abs().max().item() <= 2e-4                 # x is on a 0..255 / 58 scale
    for got, want in ((ga1, ga0), (gv1, gv0)):
        rel = ((got - want).norm() / want.norm().clamp_min(1e-12)).item()
        assert rel <= 2e-3, rel


def test_glue_kernels_alone_match_autograd_of_the_reference_expressions():
    """paif_glue_forward / paif_glue_backward against autograd of the reference expressions in fp64, including the
    evenly shared min / max gradients and the clamp masks."""
    B, H, W = 2, 37, 53
    g = torch.Generator().manual_seed(5)
    fused = (torch.rand(B, 1, H, W, generator=g) * 1.2 - 0.1).to(DEV)
    vis = (torch.rand(B, 3, H, W, generator=g) * 1.4 - 0.2).to(DEV)
    cot = torch.randn(B, 3, H, W, generator=g).to(DEV)
    for per_sample in (False, True):
        f, v = fused.clone().requires_grad_(True), vis.clone().requires_grad_(True)
        x = _GlueFn.apply(f, v, per_sample)
        x.backward(cot)
        fd, vd = fused.double().requires_grad_(True), vis.double().requires_grad_(True)
        ycc = FusionSegTask.rgb_to_ycrcb(vd)
        rgb = FusionSegTask.ycrcb_to_rgb(torch.cat([fd, ycc[:, 1:3]], 1))
        rgb = torch.where(rgb > 1, torch.ones_like(rgb), rgb)
        rgb = torch.where(rgb < 0, torch.zeros_like(rgb), rgb)
        if per_sample:
            lo, hi = rgb.amin((1, 2, 3), keepdim=True), rgb.amax((1, 2, 3), keepdim=True)
        else:
            lo, hi = rgb.min(), rgb.max()
        t = (rgb - lo) / (hi - lo)
        mean = torch.tensor([123.675, 116.28, 103.53], device=DEV, dtype=torch.float64).view(1, 3, 1, 1)
        std = torch.tensor([58.395, 57.12, 57.375], device=DEV, dtype=torch.float64).view(1, 3, 1, 1)
        ref = (t * 255 - mean) / std
        ref.backward(cot.double())
        assert (x.double() - ref).abs().max().item() <= 1e-4
        for got, want in ((f.grad, fd.grad), (v.grad, vd.grad)):
            rel = ((got.double() - want).norm() / want.norm()).item()
            assert rel <= 1e-4, (per_sample, rel)
