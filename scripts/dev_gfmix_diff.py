"""Development: where does a variant build of the fused GF kernel (PAIF_B200_LIB) differ from the two-kernel path?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paif_b200 import _lib, fusion
DEV = "cuda:0"
B, H, W = (int(a) for a in sys.argv[1:4])
torch.manual_seed(0)
z = torch.rand(B, 8, H, W, 4, device=DEV)
g = (z.amax((1, 4)) - z.amin((1, 4))).contiguous()
w = torch.randn(32, 128, 1, 1, device=DEV) * 0.15
bias = torch.randn(32, device=DEV) * 0.1
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
stats = torch.empty(3, B, H, W, device=DEV)
_lib.call("paif_gf_guide_stats", g.data_ptr(), stats.data_ptr(), B, H, W, st)
wp = fusion._pack_gf_mix(w)
lf1, lf2 = torch.empty_like(z), torch.empty_like(z)
rt = fusion._Runtime(B, H, W, 32, torch.device(DEV), _lib.ENGINE_DIRECT, False)
cw = fusion._ConvW(fusion._fold_decomp_1x1(w), 3, 1, 1)
_lib.call("paif_gf_decomp_forward", z.data_ptr(), g.data_ptr(), stats.data_ptr(), lf1.data_ptr(), lf2.data_ptr(), 32, B, H, W, st)
ref = rt.conv([lf1, lf2, z], cw, ch_shift=bias)[0]
for trial in range(3):
    out = torch.full_like(z, float("nan"))
    _lib.call("paif_gf_mix_forward", z.data_ptr(), g.data_ptr(), stats.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr(), 0, 32, B, H, W, st)
    torch.cuda.synchronize()
    err = (out - ref).abs()                                  # [B,8,H,W,4]
    bad = err > 5e-3
    print("trial", trial, "max", err.max().item(), "nan", int(torch.isnan(out).sum()), "bad px", int(bad.any(-1).any(1).sum()))
    if bad.any():
        rows = bad.any(-1).any(1).any(-1)[0].nonzero().flatten().tolist()
        cols = bad.any(-1).any(1).any(1)[0].nonzero().flatten().tolist()
        quads = bad.any(-1).any(-1).any(-1)[0].nonzero().flatten().tolist()
        print("  rows", rows[:40], "...", len(rows)); print("  cols", cols[:40], "...", len(cols)); print("  quads", quads)
