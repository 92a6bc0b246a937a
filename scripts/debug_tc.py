"""Dev tool: tcgen05 conv engine vs the direct fp32 engine and torch on the GPU (prints errors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from paif_b200 import _lib, fusion

DEV = "cuda:0"
def to_c4(t):
    B, C, H, W = t.shape
    return t.reshape(B, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()
def from_c4(t):
    B, Q, H, W, _ = t.shape
    return t.permute(0, 1, 4, 2, 3).reshape(B, Q * 4, H, W)

cases = [(1, 20, 128, 3, 1, 1), (2, 37, 200, 3, 1, 1), (1, 40, 300, 3, 1, 3), (1, 33, 130, 3, 2, 1),
         (1, 40, 256, 7, 1, 1), (2, 21, 139, 1, 1, 3), (1, 70, 640, 3, 1, 2)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (B, H, W, k, dil, nsrc) in cases:
    torch.manual_seed(0)
    xs = [torch.randn(B, 32, H, W, device=DEV) for _ in range(nsrc)]
    w = torch.randn(32, 32 * nsrc, k, k, device=DEV) * 0.1
    ref = F.conv2d(torch.cat(xs, 1).double(), w.double(), None, 1, dil * (k - 1) // 2, dil).float()
    cw = fusion._ConvW(w, nsrc, k, dil)
    assert cw.mma is not None
    outs = {}
    for name, eng in (("direct", _lib.ENGINE_DIRECT), ("tcgen05", _lib.ENGINE_TCGEN05)):
        rt = fusion._Runtime(B, H, W, 32, torch.device(DEV), eng, False)
        out, _, _, parts = rt.conv([to_c4(x) for x in xs], cw, want_partials=True)
        torch.cuda.synchronize()
        outs[name] = from_c4(out)
        err = (outs[name] - ref).abs().max().item()
        perr = (parts.sum(1) - ref.sum((2, 3))).abs().max().item()
        print("B%d %dx%d k%d d%d nsrc%d %-8s max-abs err %.3e (ref max %.2f) partial-sum err %.3e" %
              (B, H, W, k, dil, nsrc, name, err, ref.abs().max().item(), perr), flush=True)
