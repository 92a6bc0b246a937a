"""``ConvModule`` = conv(bias=False) + BN + ReLU with sub-module names ``conv`` / ``bn``
— the only configuration the reference constructs (core/segformer_head.py:50-55)."""
import torch.nn as nn


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), **kw):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding,
                              bias=norm_cfg is None)
        self.bn = nn.BatchNorm2d(out_channels) if norm_cfg is not None else None
        self.activate = nn.ReLU(inplace=True) if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


class DepthwiseSeparableConvModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("not constructed by the reference scripts")
