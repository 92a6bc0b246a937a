"""Development tool: time one conv shape with a given library build (PAIF_B200_LIB=...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paif_b200 import _lib, fusion
DEV = torch.device("cuda:0")
B, H, W = 16, 480, 640
rt = fusion._Runtime(B, H, W, 32, DEV, _lib.ENGINE_TCGEN05, False)
torch.manual_seed(0)
maps = [torch.randn(B, 8, H, W, 4, device=DEV) for _ in range(4)]
a = torch.tensor([0.25], device=DEV)
for name, nsrc, k, kw in [("k3 cin32 prelu", 1, 3, dict(slope=a)), ("k3 cin96 prelu", 3, 3, dict(slope=a)), ("k1 cin96", 3, 1, dict())]:
    w = torch.randn(32, 32 * nsrc, k, k, device=DEV) * 0.05
    cw = fusion._ConvW(w, nsrc, k, 1)
    for _ in range(3):
        rt.conv(maps[:nsrc], cw, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rt.conv(maps[:nsrc], cw, **kw)
    e1.record(); torch.cuda.synchronize()
    print("%-16s %-40s %.3f ms" % (name, os.path.basename(os.environ.get("PAIF_B200_LIB", "default")), e0.elapsed_time(e1) / 5), flush=True)
