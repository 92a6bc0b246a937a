"""Box filter by double cumulative sum with clipped borders — restatement of the
published DeepGuidedFilter ``BoxFilter`` (see package docstring)."""
import torch
from torch import nn


def diff_x(c, r):
    # c = cumsum along dim 2; window sum of radius r, clipped at both ends.
    assert c.dim() == 4
    left = c[:, :, r:2 * r + 1]
    middle = c[:, :, 2 * r + 1:] - c[:, :, :-2 * r - 1]
    right = c[:, :, -1:] - c[:, :, -2 * r - 1:-r - 1]
    return torch.cat([left, middle, right], dim=2)


def diff_y(c, r):
    assert c.dim() == 4
    left = c[:, :, :, r:2 * r + 1]
    middle = c[:, :, :, 2 * r + 1:] - c[:, :, :, :-2 * r - 1]
    right = c[:, :, :, -1:] - c[:, :, :, -2 * r - 1:-r - 1]
    return torch.cat([left, middle, right], dim=3)


class BoxFilter(nn.Module):
    def __init__(self, r):
        super().__init__()
        self.r = r

    def forward(self, x):
        assert x.dim() == 4
        return diff_y(diff_x(x.cumsum(dim=2), self.r).cumsum(dim=3), self.r)
