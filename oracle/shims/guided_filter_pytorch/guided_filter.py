"""``GuidedFilter(r, eps)(x, y)`` — restatement of the published DeepGuidedFilter
module (guide x, source y); see package docstring.  Called by the reference as
``GuidedFilter(4, eps)(res, x)`` at core/model_fusion_auto.py:529-530."""
import torch
from torch import nn

from .box_filter import BoxFilter


class GuidedFilter(nn.Module):
    def __init__(self, r, eps=1e-8):
        super().__init__()
        self.r = r
        self.eps = eps
        self.boxfilter = BoxFilter(r)

    def forward(self, x, y):
        n_x, c_x, h_x, w_x = x.size()
        n_y, c_y, h_y, w_y = y.size()
        assert n_x == n_y
        assert c_x == 1 or c_x == c_y
        assert h_x == h_y and w_x == w_y
        assert h_x > 2 * self.r + 1 and w_x > 2 * self.r + 1
        N = self.boxfilter(x.new_ones((1, 1, h_x, w_x)))
        mean_x = self.boxfilter(x) / N
        mean_y = self.boxfilter(y) / N
        cov_xy = self.boxfilter(x * y) / N - mean_x * mean_y
        var_x = self.boxfilter(x * x) / N - mean_x * mean_x
        A = cov_xy / (var_x + self.eps)
        b = mean_y - A * mean_x
        mean_A = self.boxfilter(A) / N
        mean_b = self.boxfilter(b) / N
        return mean_A * x + mean_b
