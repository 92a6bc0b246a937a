"""Shared helpers for the test-suite (imported as a top-level module: pytest puts tests/ on sys.path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["seed0_default_2x40x56", "seed1_random_1x48x72", "seed1_random_smooth_2x33x47"]


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def strided_vis(vis):
    """The channel-last strided view the reference wrappers hand to forward():
    shape [B,C,H,W], strides (C*H*W, 1, C*W, C)."""
    return vis.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
