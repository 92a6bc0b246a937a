"""Import shim (test infrastructure).  The reference does
``from antialias import Downsample as downsamp`` (operations_m.py:4); the class is
only used by the dead ``ResidualDownSample`` (operations_m.py:214), never on the
fusion hot path, so a stub that refuses to be called is sufficient."""
import torch.nn as nn


class Downsample(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        raise NotImplementedError("antialias.Downsample is off the PAIF hot path")
