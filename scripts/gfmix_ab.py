"""Development A/B of paif_gf_mix_forward builds: every library named on the command line (built by
scripts/build_gfmix_variant.sh, i.e. with -DPAIF_TC_PROFILE; e.g. one from `git stash` / an older commit as the base)
against the first one — bit-identity of the fp32 and bf16 outputs at several shapes (edge strips,
odd row counts, small batches), time at the bench shape and the role timeline (wait share per role)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paif_b200 import _lib, fusion

DEV = "cuda:0"
libs = [(os.path.basename(p), C.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ARGS = [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p]
for _, l in libs:
    l.paif_gf_mix_forward.argtypes = ARGS
    l.paif_gf_mix_forward.restype = C.c_int


def case(B, H, W, seed=0):
    torch.manual_seed(seed)
    z = torch.rand(B, 8, H, W, 4, device=DEV)
    g = (z.amax((1, 4)) - z.amin((1, 4))).contiguous()
    w = torch.randn(32, 128, 1, 1, device=DEV) * 0.15
    bias = torch.randn(32, device=DEV) * 0.1
    stats = torch.empty(3, B, H, W, device=DEV)
    _lib.call("paif_gf_guide_stats", g.data_ptr(), stats.data_ptr(), B, H, W, st)
    return z, g, stats, fusion._pack_gf_mix(w), bias


def launch(l, c, out, bf):
    z, g, stats, wp, bias = c
    B, _, H, W, _ = z.shape
    return l.paif_gf_mix_forward(z.data_ptr(), g.data_ptr(), stats.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                 int(bf), 32, B, H, W, st)


def new_out(c, bf):
    z = c[0]
    B, _, H, W, _ = z.shape
    return torch.empty(B, 4, H, W, 8, device=DEV, dtype=torch.bfloat16) if bf else torch.empty_like(z)


def run(l, c, bf):
    out = new_out(c, bf)
    out.fill_(7.0)
    rc = launch(l, c, out, bf)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out


for shape in ((3, 77, 100), (1, 480, 640), (2, 131, 52), (5, 480, 640), (2, 768, 1024), (1, 33, 12)):
    c = case(*shape)
    for bf in (False, True):
        ref = run(libs[0][1], c, bf)
        for name, l in libs[1:]:
            got = run(l, c, bf)
            same = torch.equal(ref, got)
            print("%-16s %-14s bf16=%d  bit-identical=%s%s" % (name, shape, bf, same,
                  "" if same else "  max-abs %.3e" % (ref.float() - got.float()).abs().max().item()))

c = case(16, 480, 640)
for name, l in libs:
    for bf in (False, True):
        out = new_out(c, bf)
        for _ in range(3):
            launch(l, c, out, bf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            launch(l, c, out, bf)
        e1.record()
        torch.cuda.synchronize()
        print("%-16s 16x480x640 bf16=%d  %.4f ms" % (name, bf, e0.elapsed_time(e1) / 20))
    if hasattr(l, "paif_debug_gx_counters"):
        buf = (C.c_ulonglong * 16)()
        l.paif_debug_gx_counters(buf, 1)
        run(l, c, False)
        l.paif_debug_gx_counters(buf, 0)
        print("   " + "  ".join("%s waits %4.1f%% life %.0f" % (nm, 100.0 * buf[2 * i] / max(buf[2 * i + 1], 1), buf[2 * i + 1] / (148.0 * nw))
                                for i, (nm, nw) in enumerate((("L1", 4), ("L2", 4), ("EP", 4), ("MMA", 1), ("PROD", 1)))))
