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
