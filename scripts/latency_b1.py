"""Batch-1 latency of the drop-in at 480x640 (what the PGD loop pays per frame): forward under no_grad and
forward + backward-to-input, through the public module call; native (whole-network C-ABI calls) vs per-operator path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
net, _ = bench.synth_state()
net = net.to(dev).eval()
for B in (1, 8):
    ir, vis = bench.synth_inputs(B, 480, 640, seed=1)
    ir, vis = ir.to(dev), vis.to(dev)
    gout = torch.rand(B, 1, 480, 640, device=dev) - 0.5

    def fwd():
        with torch.no_grad():
            net(ir, vis)

    def fb():
        a = ir.detach().requires_grad_(True)
        v = vis.detach().requires_grad_(True)
        net(a, v).backward(gout)

    for native in (True, False):
        net.native_forward = native
        for name, fn in (("forward", fwd), ("forward+backward", fb)):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 30
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print("batch %d  %-17s %-12s %.3f ms" % (B, name, "native" if native else "per-operator", e0.elapsed_time(e1) / n), flush=True)
