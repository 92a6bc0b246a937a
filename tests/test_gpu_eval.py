"""GPU: confusion counting kernel through ConfusionMeter, and sharding invariance of the PGD robust
evaluation (the N-rank result equals the 1-rank result bit for bit)."""
import pytest
import torch
import torch.nn as nn

import paif_b200
from oracle import fusion_oracle as fo
from paif_b200 import evaluate as ev

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class TinyTask(nn.Module):
    """Stand-in for Network_MM_Searched: fusion drop-in + a small stock-PyTorch segmentation head."""

    def __init__(self):
        super().__init__()
        self.enhance_net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
        self.head = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, stride=2), nn.ReLU(), nn.Conv2d(8, 9, 3, padding=1, stride=2))

    def forward(self, ir, vis):
        fused = self.enhance_net(ir, vis)
        x = torch.cat([fused, vis[:, 1:3]], 1)
        return fused, self.head(x)


def test_confusion_meter_matches_oracle():
    g = torch.Generator().manual_seed(0)
    label = torch.randint(0, 10, (3, 40, 56), generator=g)
    label[label == 9] = 255
    pred = torch.randint(0, 9, (3, 40, 56), generator=g)
    m = ev.ConfusionMeter(9, DEV)
    m.update(label[:2].to(DEV), pred[:2].to(DEV)).update(label[2:].to(DEV), pred[2:].to(DEV))
    assert torch.equal(m.conf.cpu(), fo.confusion_matrix(label, pred, 9))


def test_pgd_robust_eval_is_sharding_invariant():
    torch.manual_seed(0)
    model = TinyTask().to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    frames = [(torch.rand(3, 24, 40, generator=g), torch.rand(1, 24, 40, generator=g),
               torch.randint(0, 9, (24, 40), generator=g)) for _ in range(5)]
    whole = ev.robust_eval(model, frames, attack_iters=2, rank=0, world_size=1).conf.cpu()
    parts = sum(ev.robust_eval(model, frames, attack_iters=2, rank=r, world_size=3).conf.cpu() for r in range(3))
    assert torch.equal(whole, parts)
    assert int(whole.sum()) == 5 * 24 * 40
    clean = ev.robust_eval(model, frames, attack_iters=0).conf.cpu()
    assert int(clean.sum()) == 5 * 24 * 40


@pytest.mark.parametrize("n", [1, 7, 4096, 3 * 480 * 640 + 3])
def test_pgd_step_kernel_is_the_reference_update_bit_for_bit(n):
    """paif_pgd_step == attack/attack.py:504-512 (sign, step, three clamps) on the same inputs, exactly: zeros of the
    gradient do not move delta, the epsilon ball and the [0,1] box are enforced in the reference's order."""
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, generator=g).to(DEV)
    eps, alpha = 8 / 255., 2 / 255.
    delta = ((torch.rand(n, generator=g) * 2 - 1) * eps).to(DEV).requires_grad_(True)
    grad = torch.randn(n, generator=g).to(DEV)
    grad[::5] = 0.0
    delta.grad = grad.clone()
    d = torch.clamp(delta.data + alpha * torch.sign(delta.grad.data), min=-eps, max=eps)
    d = torch.clamp(d, min=-eps, max=eps)
    want = torch.clamp(d, min=0 - x, max=1 - x)
    ev.pgd_step_(delta, x, alpha, eps)
    assert torch.equal(delta.data, want)
    assert torch.equal(delta.grad, grad)                        # the running gradient sum is left to autograd


@pytest.mark.parametrize("shape", [(2, 9, 30, 40, 120, 160), (1, 9, 12, 16, 47, 61), (3, 5, 24, 40, 24, 40)])
def test_fused_loss_head_matches_interpolate_plus_cross_entropy(shape):
    """paif_segloss_forward / _backward == F.interpolate(bilinear, align_corners=False) + cross entropy(ignore_index)
    (attack/attack.py:446-448, Seg_loss :103-114), value and gradient; the gradient is reproducible bit for bit."""
    import torch.nn.functional as F
    B, K, h, w, H, W = shape
    g = torch.Generator().manual_seed(B * 100 + h)
    seg = (torch.randn(B, K, h, w, generator=g) * 2).to(DEV)
    label = torch.randint(0, K, (B, H, W), generator=g)
    label[torch.rand(B, H, W, generator=g) < 0.1] = 255
    label = label.to(DEV)
    a = seg.clone().requires_grad_(True)
    la = ev._seg_loss(a, label, 255)
    (3.0 * la).backward()
    b = seg.double().clone().requires_grad_(True)
    up = F.interpolate(b, size=(H, W), mode="bilinear", align_corners=False)
    lb = F.cross_entropy(up, label, ignore_index=255, reduction="sum") / float(H * W)
    (3.0 * lb).backward()
    assert abs(la.item() - lb.item()) <= 1e-5 * max(1.0, abs(lb.item()))
    assert (a.grad.double() - b.grad).abs().max().item() <= 1e-6 + 1e-5 * b.grad.abs().max().item()
    c = seg.clone().requires_grad_(True)
    (3.0 * ev._seg_loss(c, label, 255)).backward()
    assert torch.equal(a.grad, c.grad)


class TinyBatchTask(nn.Module):
    """Per-sample task model (what a micro-batching harness needs) with a convolutional head: every stock op on its
    path has a deterministic backward, so PGD through it is reproducible bit for bit."""

    def __init__(self):
        super().__init__()
        self.enhance_net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
        self.head = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 9, 3, padding=1))

    def forward(self, ir, vis):
        fused = self.enhance_net(ir, vis)
        x = torch.cat([fused, vis[:, 1:3]], 1)
        x = x - x.amin((1, 2, 3), keepdim=True)                 # per-sample normalisation, like FusionSegTask(per_sample_minmax)
        return fused, self.head(x)


def test_micro_batched_robust_eval_is_sharding_invariant():
    """Frames attacked in fixed-size micro-batches (SURVEY 8e caveat 1): the confusion matrix is bit-identical between
    1 rank and 2 / 3 ranks (different batch-mates, different batch positions, padded tails), with and without
    CUDA-graph replay of the PGD iteration.  Everything of ours on that path is deterministic and batch-position
    invariant (fusion kernels forward and backward, seeded starts, the fused PGD step, the integer confusion count)."""
    torch.manual_seed(0)
    model = TinyBatchTask().to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    frames = [(torch.rand(3, 40, 56, generator=g), torch.rand(1, 40, 56, generator=g),
               torch.randint(0, 9, (40, 56), generator=g)) for _ in range(7)]
    for graphed in (False, True):
        kw = dict(attack_iters=2, micro_batch=3, use_cuda_graph=graphed)
        whole = ev.robust_eval(model, frames, rank=0, world_size=1, **kw).conf.cpu()
        for world in (2, 3):
            parts = sum(ev.robust_eval(model, frames, rank=r, world_size=world, **kw).conf.cpu() for r in range(world))
            assert torch.equal(whole, parts), (graphed, world)
        assert int(whole.sum()) == 7 * 40 * 56                  # padded tail frames carry only ignore labels


def test_micro_batched_robust_eval_through_a_stock_transformer_consumer():
    """The same through the per-sample min-max wrapper and a stock-PyTorch SegFormer-shaped consumer.  Its decode head
    up-samples with bilinear interpolation, whose stock CUDA backward accumulates with float atomics: PGD steps along
    sign(grad), so a few pixels per frame can step the other way from run to run.  The all-reduced matrices therefore
    agree in total and to within a small fraction of the pixels, not bit for bit — a property of the stock consumer,
    not of the sharding (the integer all-reduce itself is exact, tests/test_eval_host.py)."""
    from paif_b200.consumer import FusionSegTask, SegFormerLite
    torch.manual_seed(0)
    fusion_net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
    seg = SegFormerLite(9, 64, dims=(16, 32, 48, 64), depths=(1, 1, 1, 1), heads=(1, 2, 3, 4))
    model = FusionSegTask(fusion_net, seg, per_sample_minmax=True).to(DEV).eval()
    for prm in seg.parameters():
        prm.requires_grad_(False)
    g = torch.Generator().manual_seed(1)
    frames = [(torch.rand(3, 64, 96, generator=g), torch.rand(1, 64, 96, generator=g),
               torch.randint(0, 9, (64, 96), generator=g)) for _ in range(7)]
    kw = dict(attack_iters=2, micro_batch=3, use_cuda_graph=True)
    whole = ev.robust_eval(model, frames, rank=0, world_size=1, **kw).conf.cpu()
    parts = sum(ev.robust_eval(model, frames, rank=r, world_size=2, **kw).conf.cpu() for r in range(2))
    assert int(whole.sum()) == int(parts.sum()) == 7 * 64 * 96
    assert int((whole - parts).abs().sum()) <= 2 * 7 * 64 * 96 // 200          # <= 0.5 % of the pixels moved


def test_fusion_cuda_graph_replay_is_bit_identical_to_eager():
    """Forward + backward-to-input of the drop-in captured in a CUDA graph and replayed returns the same bits as the
    eager call (caller-owned buffers, no hidden sync or allocation inside the library, deterministic kernels)."""
    torch.manual_seed(0)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    vis, ir = torch.rand(1, 3, 24, 40, generator=g).to(DEV), torch.rand(1, 1, 24, 40, generator=g).to(DEV)
    cot = torch.randn(1, 1, 24, 40, generator=g).to(DEV)
    a, v = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
    out = net(a, v)
    out.backward(cot)
    sa, sv = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            net(sa, sv).backward(cot)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        so = net(sa, sv)
        so.backward(cot)
    for _ in range(2):
        sa.grad.zero_()
        sv.grad.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(so, out.detach()) and torch.equal(sa.grad, a.grad) and torch.equal(sv.grad, v.grad)


def test_cuda_graph_pgd_matches_eager_pgd():
    """The graph-replayed PGD loop runs the same kernels in the same order as the eager loop.  The fusion part is
    bit-identical (previous test); the stock PyTorch head is not (bilinear-upsample backward accumulates with atomics),
    and PGD steps along sign(grad), so a pixel whose accumulated gradient is ~0 may step the other way: the two
    perturbations agree except on a small fraction of pixels, which then differ by a multiple of alpha."""
    torch.manual_seed(0)
    model = TinyTask().to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    vis, ir = torch.rand(1, 3, 24, 40, generator=g).to(DEV), torch.rand(1, 1, 24, 40, generator=g).to(DEV)
    label = torch.randint(0, 9, (1, 24, 40), generator=g).to(DEV)
    e_ir, e_vis = ev.pgd_attack_both(model, vis, ir, label, attack_iters=3, seed=5, global_index=2)
    assert e_ir.shape == ir.shape and e_vis.shape == vis.shape      # (delta_ir, delta_vis): attack/attack.py:514
    runner = ev.GraphedPGD(model, vis.shape, ir.shape, label.shape, torch.device(DEV), 8 / 255., 2 / 255.)
    for _ in range(2):                                   # twice: state is reset per frame
        g_ir, g_vis = runner.attack(vis, ir, label, 3, seed=5, global_index=2)
        for got, want in ((g_vis, e_vis), (g_ir, e_ir)):
            diff = (got - want).abs()
            assert int((diff > 1e-6).sum()) <= got.numel() // 50, int((diff > 1e-6).sum())
            assert diff.max().item() <= 2 * 3 * 2 / 255. + 1e-6          # at most every step flipped
    frames = [(vis[0].cpu(), ir[0].cpu(), label[0].cpu())] * 2
    a = ev.robust_eval(model, frames, attack_iters=2).conf.cpu()
    b = ev.robust_eval(model, frames, attack_iters=2, use_cuda_graph=True).conf.cpu()
    assert int(a.sum()) == int(b.sum()) and int((a - b).abs().sum()) <= 2 * 24 * 40 // 25
