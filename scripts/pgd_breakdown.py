"""Where one PGD iteration of the bench's PGD leg spends its time (micro-batch 8 x 480x640, reference WeTr('mit_b3') consumer
under bf16 autocast): fusion net, colour / normalisation glue, stock consumer, loss head, delta update — each forward +
backward timed alone with CUDA events, next to the whole graph-replayed iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from paif_b200.consumer import FusionSegTask, _GlueFn
from paif_b200 import evaluate as ev

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
net, _ = bench.synth_state()
net = net.to(dev).eval()
seg, seg_name = bench.build_consumer(dev)
task = FusionSegTask(net, seg, consumer_autocast=torch.bfloat16, per_sample_minmax=True, fused_glue=True).to(dev).eval()
B, H, W = 8, 480, 640
g = torch.Generator(device=dev).manual_seed(0)
vis = torch.rand(B, 3, H, W, device=dev, generator=g)
ir = torch.rand(B, 1, H, W, device=dev, generator=g)
label = torch.randint(0, 9, (B, H, W), device=dev, generator=g)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fusion_fb():
    a, v = ir.detach().requires_grad_(True), vis.detach().requires_grad_(True)
    net.forward_rgb(a, v).backward(gout1)


gout1 = torch.rand(B, 1, H, W, device=dev) - 0.5
with torch.no_grad():
    fused = net.forward_rgb(ir, vis)
    x = _GlueFn.apply(fused, vis, True)


def glue_fb():
    f = fused.detach().requires_grad_(True)
    v = vis.detach().requires_grad_(True)
    _GlueFn.apply(f, v, True).backward(gx)


gx = torch.rand_like(x) - 0.5


def consumer_fb():
    xx = x.detach().requires_grad_(True)
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        s = seg(xx)
    s.float().backward(gseg)


with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
    s0 = seg(x).float()
gseg = torch.rand_like(s0) - 0.5


def loss_fb():
    s = s0.detach().requires_grad_(True)
    ev._seg_loss(s, label, 255).backward()


d = torch.zeros_like(vis)
gd = torch.rand_like(vis) - 0.5


def step():
    d.grad = gd
    ev.pgd_step_(d, vis, 2 / 255., 8 / 255.)


def whole_eager():
    a, v = ir.detach().requires_grad_(True), vis.detach().requires_grad_(True)
    _, s = task(a, v)
    ev._seg_loss(s, label, 255).backward()


print("micro-batch %d x %dx%d, consumer: %s" % (B, H, W, seg_name))
for name, fn in (("fusion net fwd+bwd (forward_rgb)", fusion_fb), ("colour / normalisation glue fwd+bwd", glue_fb),
                 ("stock consumer fwd+bwd (bf16 autocast)", consumer_fb), ("loss head fwd+bwd", loss_fb),
                 ("delta update (one modality)", step), ("whole iteration, eager", whole_eager)):
    print("%-42s %8.2f ms" % (name, timeit(fn)), flush=True)
runner = ev.GraphedPGD(task, vis.shape, ir.shape, label.shape, dev, 8 / 255., 2 / 255.)
torch.cuda.synchronize()
print("%-42s %8.2f ms" % ("whole iteration, CUDA-graph replay", timeit(lambda: runner.graph.replay())), flush=True)
