"""Development: time paif_gf_mix_forward against the two-kernel path (guided filter + 1x1 on the engine)."""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paif_b200 import _lib, fusion

DEV = "cuda:0"
B, H, W = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 480, 640)))
torch.manual_seed(0)
z = torch.rand(B, 8, H, W, 4, device=DEV)
g = (z.amax((1, 4)) - z.amin((1, 4))).contiguous()
w = torch.randn(32, 128, 1, 1, device=DEV) * 0.15
bias = torch.randn(32, device=DEV) * 0.1
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
stats = torch.empty(3, B, H, W, device=DEV)
_lib.call("paif_gf_guide_stats", g.data_ptr(), stats.data_ptr(), B, H, W, st)
wp = fusion._pack_gf_mix(w)
out = torch.empty_like(z)
out16 = torch.empty(B, 4, H, W, 8, device=DEV, dtype=torch.bfloat16)
lf1, lf2 = torch.empty_like(z), torch.empty_like(z)
rt = fusion._Runtime(B, H, W, 32, torch.device(DEV), _lib.ENGINE_TCGEN05, False)
cw = fusion._ConvW(fusion._fold_decomp_1x1(w), 3, 1, 1)


def new(bf=False):
    _lib.call("paif_gf_mix_forward", z.data_ptr(), g.data_ptr(), stats.data_ptr(), wp.data_ptr(), bias.data_ptr(),
              (out16 if bf else out).data_ptr(), int(bf), 32, B, H, W, st)


def old():
    _lib.call("paif_gf_decomp_forward", z.data_ptr(), g.data_ptr(), stats.data_ptr(), lf1.data_ptr(), lf2.data_ptr(), 32, B, H, W, st)
    return rt.conv([lf1, lf2, z], cw, ch_shift=bias)[0]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


new()
ref = old()
torch.cuda.synchronize()
print("max-abs new vs old (TF32 both): %.3e   max|ref| %.3f" % ((out - ref).abs().max().item(), ref.abs().max().item()))
t_new, t_new16, t_old = timeit(new), timeit(lambda: new(True)), timeit(old)
maps = B * H * W * 128 / 1e9
print("B=%d %dx%d: fused %.3f ms (%.0f GB/s on 2 maps), fused->bf16 %.3f ms, gf + 1x1 %.3f ms" %
      (B, H, W, t_new, 2 * maps / t_new * 1e3, t_new16, t_old))

if os.environ.get("PAIF_B200_PROFILE_LIB") == "1":
    lib = _lib.load()
    buf = (ctypes.c_ulonglong * 16)()
    lib.paif_debug_gx_counters(buf, 1)
    new()
    lib.paif_debug_gx_counters(buf, 0)
    for i, (name, nw) in enumerate((("L1", 4), ("L2", 4), ("EP", 4), ("MMA", 1))):
        wait, tot = buf[2 * i], buf[2 * i + 1]
        print("%-4s waits %5.1f%% of its life (avg life %.0f cycles per warp)" % (name, 100.0 * wait / max(tot, 1), tot / (148.0 * nw)))
