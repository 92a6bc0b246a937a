"""TEST INFRASTRUCTURE ONLY.  Imports the *unmodified* reference from
``/root/reference`` with the import shims in ``oracle/shims`` so that the restatement
in ``fusion_oracle.py`` can be validated against the real thing and golden vectors can
be generated (``gen_golden.py``).  ``/root/reference`` exists only in the build
container — nothing that runs on the GPU box may call this module.
"""
import contextlib
import io
import os
import sys
from collections import namedtuple

_HERE = os.path.dirname(os.path.abspath(__file__))
#: where build() stages a copy of the reference tree for the GPU box (git-ignored, shipped by gpurun)
STAGED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def reference_root():
    """PAIF_REFERENCE_ROOT, else /root/reference (build container), else the staged copy baseline/_ref."""
    env = os.environ.get("PAIF_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", STAGED_ROOT):
        if os.path.isfile(os.path.join(cand, "core", "model_fusion_auto.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = reference_root()
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

# test_original.py:709-713 == robust_test.py:253-257 (the only shipped genotype)
Genotype = namedtuple('Genotype', 'normal_1 normal_1_concat normal_2 normal_2_concat normal_3 normal_3_concat')
fusion_at = Genotype(normal_1=[('Denseblocks_3_1', 0), ('DilConv_3_2', 1)], normal_1_concat=[1, 2],
                     normal_2=[('Denseblocks_3_1', 0), ('Denseblocks_3_1', 1)], normal_2_concat=[1, 2],
                     normal_3=[('ECAattention_3', 0), ('Residualblocks_7_1', 1)], normal_3_concat=[1, 2])


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "model_fusion_auto.py"))


def install_shims():
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)


def load_reference():
    """Return the reference's ``core.model_fusion_auto`` module (shims installed)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    import core.model_fusion_auto as m  # noqa: E402
    return m


def build_reference_fusion(genotype=fusion_at, C=32, seed=0, randomize=False):
    """``Network_Fusion_Searched(C, None, genotype)`` (core/model_fusion_auto.py:599)
    with PyTorch default init under ``torch.manual_seed(seed)``; ``randomize`` perturbs
    BN statistics / affine and PReLU slopes as SURVEY.md §8d's second weight set."""
    import torch
    m = load_reference()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.Network_Fusion_Searched(C, None, genotype)
    if randomize:
        randomize_state(net, seed)
    return net.eval()


def randomize_state(net, seed):
    import torch
    g = torch.Generator().manual_seed(1000 + seed)
    with torch.no_grad():
        for name, mod in net.named_modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            elif isinstance(mod, torch.nn.PReLU):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.45 + 0.05)
