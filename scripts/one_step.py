"""One step of the bench workload bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures
(scripts/prof_r2.sh).  `fwd`: one forward at batch 16 x 480x640 through the per-operator path (same kernels and launch
geometry as the single-call path `bench.py` times); `bwd`: one forward + backward-to-input step at batch 8;
`bf16`: the forward with net.storage='bf16'."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if mode == "bwd" else 16)
    H, W = 480, 640
    dev = torch.device("cuda", 0)
    net, _ = bench.synth_state()
    net = net.to(dev).eval()
    ir, vis = bench.synth_inputs(B, H, W, seed=1)
    ir, vis = ir.to(dev), vis.to(dev)
    if mode == "bf16":
        net.storage = 'bf16'
    net.native_forward = False                 # per-operator path: the same launches, one ctypes call each
    gout = torch.rand(B, 1, H, W, device=dev) - 0.5

    def step():
        if mode == "bwd":
            a = ir.detach().requires_grad_(True)
            v = vis.detach().requires_grad_(True)
            net(a, v).backward(gout)
        else:
            with torch.no_grad():
                net(ir, vis)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("one_step %s: %d launches" % (mode, net.last_launches))


if __name__ == "__main__":
    main()
