"""compute-sanitizer driver for the fused decomposition kernel (the sanitizer build: `python -m paif_b200.build --sanitize`,
PAIF_B200_LIB=paif_b200/libpaif_b200_san.so): small shapes with edge strips, several chunks per CTA, fp32 / bf16 / saving
variants; prints a checksum per launch (parity itself is what tests/test_gpu_kernels.py checks on the production build)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paif_b200 import _lib, fusion

DEV = "cuda:0"
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (B, H, W) in ((1, 33, 12), (2, 40, 56), (1, 70, 100)):
    torch.manual_seed(0)
    z = torch.rand(B, 8, H, W, 4, device=DEV)
    g = (z.amax((1, 4)) - z.amin((1, 4))).contiguous()
    stats = torch.empty(3, B, H, W, device=DEV)
    _lib.call("paif_gf_guide_stats", g.data_ptr(), stats.data_ptr(), B, H, W, st)
    wp = fusion._pack_gf_mix(torch.randn(32, 128, 1, 1, device=DEV) * 0.15)
    bias = torch.randn(32, device=DEV) * 0.1
    out = torch.empty_like(z)
    out16 = torch.empty(B, 4, H, W, 8, device=DEV, dtype=torch.bfloat16)
    ma = torch.empty_like(z)
    _lib.call("paif_gf_mix_forward", z.data_ptr(), g.data_ptr(), stats.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr(), 0, 32, B, H, W, st)
    _lib.call("paif_gf_mix_forward", z.data_ptr(), g.data_ptr(), stats.data_ptr(), wp.data_ptr(), bias.data_ptr(), out16.data_ptr(), 1, 32, B, H, W, st)
    _lib.call("paif_gf_mix_forward_save", z.data_ptr(), g.data_ptr(), stats.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr(), 0, ma.data_ptr(), 32, B, H, W, st)
    torch.cuda.synchronize()
    print((B, H, W), "sum out %.6f  out16 %.4f  mean_a %.6f" % (out.double().sum().item(), out16.double().sum().item(), ma.double().sum().item()), flush=True)
print("done")
