"""Development tool: role timeline of the tcgen05 conv engine (wait vs busy cycles per role), using the
-DPAIF_TC_PROFILE build.  python -m paif_b200.build --profile; PAIF_B200_PROFILE_LIB=1 python scripts/tc_timeline.py [--bf16]"""
import ctypes, os, sys
os.environ["PAIF_B200_PROFILE_LIB"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paif_b200 import _lib, fusion
lib = _lib.load()
lib.paif_debug_tc_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
DEV = torch.device("cuda:0")
B, H, W = 16, 480, 640
BF16 = "--bf16" in sys.argv                      # bf16 storage mode: C8 maps, kind::f16 MMAs
rt = fusion._Runtime(B, H, W, 32, DEV, _lib.ENGINE_TCGEN05, False, bf16=BF16)
buf = (ctypes.c_ulonglong * 16)()
def counters(reset=1):
    lib.paif_debug_tc_counters(buf, reset)
    return list(buf)
torch.manual_seed(0)
maps = [torch.randn(B, 4, H, W, 8, device=DEV).to(torch.bfloat16) if BF16 else torch.randn(B, 8, H, W, 4, device=DEV)
        for _ in range(6)]
a = torch.tensor([0.25], device=DEV)
cases = [("k3 cin32 prelu", 1, 3, 1, dict(slope=a)),
         ("k3 cin96 prelu+3res", 3, 3, 1, dict(slope=a, post_scale=0.333, post_res=maps[3:6])),
         ("k7 cin32", 1, 7, 1, dict()),
         ("k3d2 cin32 bn prelu 3res pre", 1, 3, 2, dict(slope=a, post_res=maps[3:6], want_pre=True)),
         ("k1 cin32 3res", 1, 1, 1, dict(post_res=maps[3:6])),
         ("k1 cin96", 3, 1, 1, dict())]
for name, nsrc, k, dil, kw in cases:
    w = torch.randn(32, 32 * nsrc, k, k, device=DEV) * 0.05
    cw = fusion._ConvW(w, nsrc, k, dil)
    for _ in range(2):
        rt.conv(maps[:nsrc], cw, **kw)
    counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rt.conv(maps[:nsrc], cw, **kw); e1.record()
    torch.cuda.synchronize()
    c = counters()
    ms = e0.elapsed_time(e1)
    def f(i):   # wait share of the role's life (summed over CTAs / segments)
        return c[2 * i] / max(c[2 * i + 1], 1)
    print("%-30s %.3f ms | epilogue: wait %.0f%% | issuer: waits for input %.0f%%, for an accumulator slot %.0f%% | producer: wait %.0f%%" % (
        name, ms, 100 * f(0), 100 * f(1), 100 * c[9] / max(c[3], 1), 100 * f(2)), flush=True)

# stem_out on the engine (single-output mode)
import paif_b200
net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
p = net._packed(False)
out = torch.empty(B, 1, H, W, device=DEV)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
args = (maps[0].data_ptr(), (p["out_mma16"] if BF16 else p["out_mma"]).data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(),
        out.data_ptr(), None, _lib.STORAGE_BF16 if BF16 else _lib.STORAGE_F32, 32, B, H, W, st)
for _ in range(2):
    _lib.call("paif_out_forward_tc", *args)
counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _lib.call("paif_out_forward_tc", *args); e1.record()
torch.cuda.synchronize()
c = counters()
print("stem_out k5 CP16: %.3f ms | epilogue wait %.0f%% of %.0f kcyc-total | mma wait %.0f%% | producer wait %.0f%%" % (
    e0.elapsed_time(e1), 100 * c[0] / max(c[1], 1), c[1] / 1e3, 100 * c[2] / max(c[3], 1), 100 * c[4] / max(c[5], 1)))
