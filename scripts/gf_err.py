"""Development check: guided-filter forward error vs the fp32 (reference algorithm) and fp64 oracles."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import fusion_oracle as fo
from paif_b200 import _lib
DEV = "cuda:0"
def to_c4(t):
    B, C, H, W = t.shape
    return t.reshape(B, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()
def from_c4(t):
    B, Q, H, W, _ = t.shape
    return t.permute(0, 1, 4, 2, 3).reshape(B, Q * 4, H, W)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (B, H, W), smooth in [((2, 40, 56), False), ((1, 200, 236), False), ((1, 200, 236), True), ((1, 480, 640), False), ((1, 480, 640), True), ((1, 768, 1024), True), ((12, 480, 640), False)]:
    torch.manual_seed(1)
    z = torch.rand(B, 32, H, W)
    if smooth:
        z = F.avg_pool2d(z, 9, 1, 4)
    res = fo.get_residue(z)
    LF, _ = fo.decomposition(z)
    LF64, _ = fo.decomposition(z.double())
    zc = to_c4(z).to(DEV)
    lf1, lf2 = torch.empty_like(zc), torch.empty_like(zc)
    resd = res[:, 0].contiguous().to(DEV)
    stats = torch.empty(3, B, H, W, device=DEV)
    _lib.call("paif_gf_guide_stats", resd.data_ptr(), stats.data_ptr(), B, H, W, st)
    _lib.call("paif_gf_decomp_forward", zc.data_ptr(), resd.data_ptr(), stats.data_ptr(), lf1.data_ptr(), lf2.data_ptr(), 32, B, H, W, st)
    got = torch.cat([from_c4(lf1), from_c4(lf2)], 1).cpu()
    e32 = (got - LF).abs().max().item()
    e64 = (got.double() - LF64).abs().max().item()
    noise = (LF.double() - LF64).abs().max().item()
    print("shape %s smooth %d: ours-vs-fp64 %.3e   ours-vs-fp32ref %.3e   fp32ref-vs-fp64 %.3e" % ((B, H, W), smooth, e64, e32, noise), flush=True)
