"""Development tool: every conv shape of the shipped genotype on the tcgen05 engine, persistent vs tiled launch
(paif_conv_set_persistent), at the bench workload.  python scripts/conv_ab.py [B H W] [--bf16]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paif_b200
from paif_b200 import _lib, fusion

DEV = torch.device("cuda:0")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, H, W = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (16, 480, 640)
BF16 = "--bf16" in sys.argv
lib = _lib.load()
rt = fusion._Runtime(B, H, W, 32, DEV, _lib.ENGINE_TCGEN05, False, bf16=BF16)
torch.manual_seed(0)
maps = [torch.randn(B, 4, H, W, 8, device=DEV).to(torch.bfloat16) if BF16 else torch.randn(B, 8, H, W, 4, device=DEV)
        for _ in range(7)]
a = torch.tensor([0.25], device=DEV)
MAP = B * H * W * 32 * (2 if BF16 else 4)
cases = [("k3 cin32 prelu            (2 maps)", 1, 3, 1, dict(slope=a), 2),
         ("k3 cin64 prelu            (3 maps)", 2, 3, 1, dict(slope=a), 3),
         ("k3 cin96 +x               (5 maps)", 3, 3, 1, dict(post_scale=0.333, post_res=maps[3:4]), 5),
         ("k3 cin96 +x +relu out     (6 maps)", 3, 3, 1, dict(post_scale=0.333, post_res=maps[3:4], act2_slope=a), 6),
         ("k3 cin96 +3 res           (7 maps)", 3, 3, 1, dict(post_scale=0.333, post_res=maps[3:6]), 7),
         ("k3d2 cin32 3res (DilConv) (5 maps)", 1, 3, 2, dict(post_res=maps[3:6]), 5),
         ("k3d2 cin32 prelu 2res     (4 maps)", 1, 3, 2, dict(slope=a, post_res=maps[3:5]), 4),
         ("k3d2 cin32 no res         (2 maps)", 1, 3, 2, dict(slope=a), 2),
         ("k3 cin32 mask + 2res bwd  (5 maps)", 1, 3, 1, dict(mask_src=maps[5], mask_slope=a, post_res=maps[3:5]), 5),
         ("k7 cin32 prelu            (2 maps)", 1, 7, 1, dict(slope=a), 2),
         ("k3 cin32 +res             (3 maps)", 1, 3, 1, dict(post_res=maps[3:4]), 3)]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("B=%d %dx%d %s" % (B, H, W, "bf16" if BF16 else "fp32"))
for name, nsrc, k, dil, kw, nmaps in cases:
    w = torch.randn(32, 32 * nsrc, k, k, device=DEV) * 0.05
    cw = fusion._ConvW(w, nsrc, k, dil)
    res = {}
    for mode in (0, 1):
        lib.paif_conv_set_persistent(mode)
        res[mode] = timeit(lambda: rt.conv(maps[:nsrc], cw, **kw))
    gbs = nmaps * MAP / 1e9
    print("%-38s tiled %.3f ms (%4.0f GB/s)   persistent %.3f ms (%4.0f GB/s)  %+.1f%%" % (
        name, res[0], gbs / res[0] * 1e3, res[1], gbs / res[1] * 1e3, 100 * (res[1] / res[0] - 1)), flush=True)

# stem_out on the engine (single-output mode)
net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
p = net._packed(False)
out = torch.empty(B, 1, H, W, device=DEV)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
oargs = (maps[0].data_ptr(), (p["out_mma16"] if BF16 else p["out_mma"]).data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(),
         out.data_ptr(), None, _lib.STORAGE_BF16 if BF16 else _lib.STORAGE_F32, 32, B, H, W, st)
res = {}
for mode in (0, 1):
    lib.paif_conv_set_persistent(mode)
    res[mode] = timeit(lambda: _lib.call("paif_out_forward_tc", *oargs))
print("%-38s tiled %.3f ms   persistent %.3f ms  %+.1f%%" % ("stem_out k5 CP16", res[0], res[1], 100 * (res[1] / res[0] - 1)))
lib.paif_conv_set_persistent(1)
