"""Sharded (one process per GPU) attack-tolerant evaluation around the fusion drop-in.

What the reference does in one process with ``batch_size=1`` (robust_test.py:95-212):
``attack_both`` PGD (attack/attack.py:417-514) -> forward on the attacked pair -> arg-max ->
``sklearn.metrics.confusion_matrix(labels=0..8)`` summed on the host -> ``compute_results``
(util/util.py:31-55).  Here the frames are split contiguously over the ranks of a
``torch.distributed`` group, every rank accumulates an int64 confusion matrix on its GPU
(``paif_confusion_accumulate``) and ONE all-reduce(SUM) of 81 int64 values merges them: integer
addition is associative, so the N-GPU matrix is bit-identical to the 1-GPU one.

Sharding invariance needs three things besides the integer reduction, all handled here:
 * the PGD start point is drawn from a generator keyed by (seed, GLOBAL frame index), not from the
   unseeded global RNG the reference uses (attack/attack.py:434);
 * frames are evaluated one at a time through the task wrapper, because its min-max normalisation
   spans the whole batch it is given (core/model_fusion_auto.py:721-723; the reference only ever
   passes batch 1);
 * the fusion kernels are batch-position invariant and deterministic (tests/test_gpu_parity.py).

The segmentation consumer is any stock-PyTorch ``nn.Module`` mapping ``[B,3,H,W] -> [B,K,h,w]``
logits (the reference uses SegFormer MiT-B3, which stays stock PyTorch by the scope contract).
"""
import collections
import ctypes

import torch
import torch.nn.functional as F

from . import _lib


def shard_range(n_items, rank, world_size):
    """Contiguous split of ``range(n_items)``: the first ``n_items % world_size`` ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class ConfusionMeter:
    """int64 confusion matrix (rows = label, cols = prediction; labels outside 0..K-1, e.g. the
    ignore index 255, are skipped exactly as sklearn's ``labels=`` argument does, robust_test.py:210)."""

    def __init__(self, num_classes=9, device="cuda"):
        self.num_classes = num_classes
        self.conf = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=device)

    def update(self, label, pred):
        """label, pred: integer CUDA tensors of equal numel."""
        if not (label.is_cuda and pred.is_cuda):
            raise RuntimeError("ConfusionMeter.update counts on the GPU: label and pred must be CUDA tensors")
        label = label.reshape(-1).to(torch.int64).contiguous()
        pred = pred.reshape(-1).to(torch.int64).contiguous()
        if label.numel() != pred.numel():
            raise ValueError("label and pred differ in size")
        with torch.cuda.device(label.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(label.device).cuda_stream)
            _lib.call("paif_confusion_accumulate", label.data_ptr(), pred.data_ptr(), label.numel(),
                      self.num_classes, self.conf.data_ptr(), stream)
        return self

    def all_reduce(self, group=None):
        """Sum the matrices of all ranks in place (NCCL over NVLink for CUDA tensors, gloo for CPU ones)."""
        all_reduce_confusion(self.conf, group)
        return self

    def results(self):
        return compute_results(self.conf)


def all_reduce_confusion(conf, group=None):
    """all-reduce(SUM) of an int64 confusion matrix; a no-op without an initialised process group."""
    import torch.distributed as dist
    if conf.dtype != torch.int64:
        raise TypeError("confusion matrices are reduced as int64 so that the sum is exact")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(conf, op=dist.ReduceOp.SUM, group=group)
    return conf


def compute_results(conf_total):
    """precision / recall / IoU per class with the reference's conventions (util/util.py:31-55:
    unlabeled class included, NaN where a denominator is zero).  Returns three float64 tensors."""
    c = conf_total.detach().to("cpu", torch.float64)
    tp = c.diag()
    col, row = c.sum(0), c.sum(1)
    nan = torch.full_like(tp, float("nan"))
    precision = torch.where(col == 0, nan, tp / col)
    recall = torch.where(row == 0, nan, tp / row)
    union = row + col - tp
    iou = torch.where(union == 0, nan, tp / union)
    return precision, recall, iou


def seeded_delta(shape, epsilon, seed, global_index, device):
    """U(-eps, eps) PGD start point that depends only on (seed, global frame index).  ``shape`` is the shape of ONE
    frame's batch-1 tensor ``[1, C, H, W]``; ``global_index`` may be a sequence, giving ``[len, C, H, W]``."""
    if isinstance(global_index, (list, tuple, range)):
        return torch.cat([seeded_delta(shape, epsilon, seed, gi, device) for gi in global_index], 0)
    g = torch.Generator(device="cpu").manual_seed((int(seed) * 1000003 + int(global_index)) % (2 ** 63 - 1))
    return ((torch.rand(shape, generator=g) * 2.0 - 1.0) * epsilon).to(device)


#: PGD perturbations in the order the reference's ``attack_both`` returns them (attack/attack.py:514:
#: ``return delta_ir, delta_vis``) — NOT (vis, ir): the two broadcast against each other's image
#: ([B,3,H,W] + [B,1,H,W]) without an error, so the order is pinned by name and by tests/test_eval_host.py.
PGDDelta = collections.namedtuple("PGDDelta", "delta_ir delta_vis")


def _replay_signature(model, vis_shape, ir_shape, label_shape, device, epsilon, alpha):
    """Everything a captured PGD iteration depends on besides the data it copies in: the shapes, the step sizes and,
    for every paif_b200 fusion net inside ``model``, the identity of its packed weights (a graph replay never
    re-enters Python, so it would keep reading the device pointers of a pack that ``load_state_dict`` or an engine
    switch has since replaced and freed)."""
    sigs = tuple(m.pack_signature(True) for m in model.modules() if hasattr(m, "pack_signature"))
    return (tuple(vis_shape), tuple(ir_shape), tuple(label_shape), str(device), float(epsilon), float(alpha), sigs)


class _SegLossFn(torch.autograd.Function):
    """Bilinear up-sampling of the logits to the label size + cross entropy (sum over the valid pixels / (H W)) as the
    kernels ``paif_segloss_forward`` / ``paif_segloss_backward``; the gradient w.r.t. the logits is a fixed-order
    gather, so it is reproducible bit for bit (the stock bilinear backward accumulates with float atomics)."""

    @staticmethod
    def forward(ctx, seg, label, ignore_index):
        B, K, h, w = seg.shape
        H, W = label.shape[1:]
        seg = seg.contiguous().float()
        label = label.contiguous().long()
        dev = seg.device
        nblk = _lib.load().paif_glue_blocks(H, W)
        partial = torch.empty((B, nblk), device=dev, dtype=torch.float32)
        gup = torch.empty((B, K, H, W), device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.call("paif_segloss_forward", seg.data_ptr(), label.data_ptr(), partial.data_ptr(),
                      None if gup is None else gup.data_ptr(), int(ignore_index), 1.0 / float(H * W), K, B, h, w, H, W, stream)
        ctx.gup, ctx.dims = gup, (K, B, h, w, H, W)
        return partial.sum()

    @staticmethod
    def backward(ctx, g):
        K, B, h, w, H, W = ctx.dims
        gup = ctx.gup
        g = g.contiguous().float()
        gseg = torch.empty((B, K, h, w), device=gup.device, dtype=torch.float32)
        with torch.cuda.device(gup.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(gup.device).cuda_stream)
            _lib.call("paif_segloss_backward", gup.data_ptr(), g.data_ptr(), gseg.data_ptr(), K, B, h, w, H, W, stream)
        return gseg, None, None


def _seg_loss(seg, label, ignore_index):
    """``Seg_loss`` of attack/attack.py:103-114 (bilinear up-sampling to the label size + cross entropy with
    ``ignore_index=255``).  The reference averages over the valid pixels of its batch-1 input; here the per-pixel
    losses are SUMMED and divided by the constant H*W, so that a frame's gradient does not depend on which other
    frames share its micro-batch (PGD steps along sign(grad): a positive constant factor changes nothing).
    CUDA logits with <= 16 classes go through the fused, deterministic loss-head kernels."""
    if seg.is_cuda and seg.shape[1] <= 16:
        return _SegLossFn.apply(seg, label, ignore_index)
    seg = F.interpolate(seg, size=label.shape[1:], mode="bilinear", align_corners=False)
    return F.cross_entropy(seg, label, ignore_index=ignore_index, reduction="sum") / float(label.shape[1] * label.shape[2])


def _frame_indices(global_index, n):
    if isinstance(global_index, (list, tuple, range)):
        if len(global_index) != n:
            raise ValueError("one global frame index per frame of the batch")
        return list(global_index)
    if n != 1:
        raise ValueError("a batch of %d frames needs %d global frame indices" % (n, n))
    return [global_index]


def pgd_step_(delta, x, alpha, epsilon):
    """In-place PGD step ``delta <- clamp(clamp(delta + alpha sign(delta.grad), +-eps), 0 - x, 1 - x)``
    (attack/attack.py:504-512) as ONE kernel (``paif_pgd_step``) instead of six elementwise launches."""
    d, g, x = delta.data, delta.grad, x.contiguous()
    if not (d.is_cuda and d.is_contiguous() and g.is_contiguous() and d.dtype == torch.float32
            and g.dtype == torch.float32 and x.dtype == torch.float32 and g.shape == d.shape == x.shape):
        raise RuntimeError("pgd_step_: contiguous fp32 CUDA tensors of one shape expected")
    with torch.cuda.device(d.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)
        _lib.call("paif_pgd_step", d.data_ptr(), g.data_ptr(), x.data_ptr(), float(alpha), float(epsilon), d.numel(), stream)
    return delta


def pgd_attack_both(model, x_vis, x_ir, label, epsilon=8 / 255., alpha=2 / 255., attack_iters=10,
                    seed=0, global_index=0, ignore_index=255):
    """``attack_both(..., attack_loss='l_seg', attack_way='PGD')`` (attack/attack.py:417-514) with a seeded start,
    for one frame or a batch of independent frames (``global_index``: one index per frame; ``model`` must not couple
    the frames of a batch — see ``FusionSegTask.per_sample_minmax``).  ``model(ir, vis) -> (fused, seg_logits)``.
    Faithful to the reference including its quirk of never zeroing ``delta.grad`` (the step uses the sign of the
    running sum of gradients, attack/attack.py:501-512).  Returns :class:`PGDDelta` ``(delta_ir, delta_vis)``, the
    reference's order."""
    dev = x_vis.device
    idx = _frame_indices(global_index, x_vis.shape[0])
    d_vis = seeded_delta((1,) + tuple(x_vis.shape[1:]), epsilon, seed, [2 * i for i in idx], dev)
    d_ir = seeded_delta((1,) + tuple(x_ir.shape[1:]), epsilon, seed, [2 * i + 1 for i in idx], dev)
    d_vis = torch.max(torch.min(d_vis, 1 - x_vis), 0 - x_vis).requires_grad_(True)
    d_ir = torch.max(torch.min(d_ir, 1 - x_ir), 0 - x_ir).requires_grad_(True)
    for _ in range(attack_iters):
        with torch.enable_grad():
            _, seg = model(x_ir + d_ir, x_vis + d_vis)
            loss = _seg_loss(seg, label, ignore_index)
        loss.backward()
        with torch.no_grad():
            for d, x in ((d_vis, x_vis), (d_ir, x_ir)):
                pgd_step_(d, x, alpha, epsilon)
    return PGDDelta(delta_ir=d_ir.detach(), delta_vis=d_vis.detach())


class GraphedPGD:
    """One PGD iteration of :func:`pgd_attack_both` (forward, loss, backward to the inputs, delta update)
    captured once in a CUDA graph and replayed ``attack_iters`` times per micro-batch.  Same arithmetic, same
    kernels; what disappears is the host time of launching ~1500 small stock-PyTorch kernels per iteration of
    the segmentation consumer (and ~75 of ours).  The fusion kernels take every buffer from the caller and never
    synchronise, so they are capturable as they are."""

    def __init__(self, model, vis_shape, ir_shape, label_shape, device, epsilon, alpha, ignore_index=255):
        self.eps, self.alpha = float(epsilon), float(alpha)
        self.x_vis = torch.zeros(vis_shape, device=device)
        self.x_ir = torch.zeros(ir_shape, device=device)
        self.label = torch.zeros(label_shape, device=device, dtype=torch.long)
        self.d_vis = torch.zeros(vis_shape, device=device, requires_grad=True)
        self.d_ir = torch.zeros(ir_shape, device=device, requires_grad=True)
        self.model, self.ignore_index = model, ignore_index
        self.signature = None
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up: workspaces, packed weights, .grad buffers
                self._iteration()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        # taken AFTER capture: the warm-up iterations are what (re)build the pack the graph points into
        self.signature = _replay_signature(model, vis_shape, ir_shape, label_shape, device, epsilon, alpha)

    def matches(self, vis_shape, ir_shape, label_shape, device, epsilon, alpha):
        """True when replaying this graph is still valid for such a call (same shapes / steps / packed weights)."""
        return self.signature == _replay_signature(self.model, vis_shape, ir_shape, label_shape, device, epsilon, alpha)

    def _iteration(self):
        with torch.enable_grad():
            _, seg = self.model(self.x_ir + self.d_ir, self.x_vis + self.d_vis)
            loss = _seg_loss(seg, self.label, self.ignore_index)
        loss.backward()
        with torch.no_grad():
            for d, x in ((self.d_vis, self.x_vis), (self.d_ir, self.x_ir)):
                pgd_step_(d, x, self.alpha, self.eps)

    def attack(self, x_vis, x_ir, label, attack_iters, seed, global_index):
        dev = self.x_vis.device
        if not self.matches(x_vis.shape, x_ir.shape, label.shape, dev, self.eps, self.alpha):
            raise RuntimeError("GraphedPGD: the captured graph no longer matches this call (input shapes changed, or "
                               "the fusion net's weights / engine switches changed after capture, so the graph "
                               "points at a stale weight pack); build a new GraphedPGD")
        idx = _frame_indices(global_index, x_vis.shape[0])
        with torch.no_grad():
            self.x_vis.copy_(x_vis)
            self.x_ir.copy_(x_ir)
            self.label.copy_(label)
            for d, x, off in ((self.d_vis, self.x_vis, 0), (self.d_ir, self.x_ir, 1)):
                init = seeded_delta((1,) + tuple(x.shape[1:]), self.eps, seed, [2 * i + off for i in idx], dev)
                d.copy_(torch.max(torch.min(init, 1 - x), 0 - x))
                d.grad.zero_()                      # a fresh delta per frame, as attack/attack.py:433-441
        for _ in range(attack_iters):
            self.graph.replay()
        return PGDDelta(delta_ir=self.d_ir.detach().clone(), delta_vis=self.d_vis.detach().clone())


def robust_eval(model, frames, num_classes=9, attack_iters=10, epsilon=8 / 255., alpha=2 / 255., seed=0,
                rank=0, world_size=1, group=None, use_cuda_graph=False, micro_batch=1, ignore_index=255):
    """Evaluate this rank's shard of ``frames`` (an indexable of ``(vis[3,H,W], ir[1,H,W], label[H,W])``
    host or device tensors) under PGD and return the all-reduced :class:`ConfusionMeter`.
    ``attack_iters=0`` gives the clean evaluation of test_original.py:98-258.

    ``micro_batch`` frames are attacked and evaluated together.  It needs a ``model`` that treats the frames of a
    batch independently (``FusionSegTask(per_sample_minmax=True)``; the reference wrapper's min-max spans the batch,
    core/model_fusion_auto.py:721-723, so with it only ``micro_batch=1`` reproduces the reference).  Every launch
    uses exactly ``micro_batch`` frames — a ragged tail is padded with copies of the shard's last frame whose labels
    are all ``ignore_index`` — so the kernels see the same shapes whatever the sharding, which is what keeps the
    all-reduced matrix bit-identical between 1 and N GPUs."""
    dev = next(model.parameters()).device
    meter = ConfusionMeter(num_classes, dev)
    runner = getattr(model, "_paif_pgd_runner", None)      # the captured graph is reused across calls
    mine = list(shard_range(len(frames), rank, world_size))
    mb = max(1, int(micro_batch))
    for lo in range(0, len(mine), mb):
        idx = mine[lo:lo + mb]
        items = [frames[gi] for gi in idx]
        pad = mb - len(idx)
        vis = torch.stack([f[0].to(dev).float() for f in items] + [items[-1][0].to(dev).float()] * pad)
        ir = torch.stack([f[1].to(dev).float() for f in items] + [items[-1][1].to(dev).float()] * pad)
        label = torch.stack([f[2].to(dev).long() for f in items] +
                            [torch.full_like(items[-1][2].to(dev).long(), ignore_index)] * pad)
        gidx = idx + [idx[-1]] * pad
        if attack_iters > 0:
            if use_cuda_graph:
                if runner is None or not runner.matches(vis.shape, ir.shape, label.shape, dev, epsilon, alpha):
                    runner = GraphedPGD(model, vis.shape, ir.shape, label.shape, dev, epsilon, alpha, ignore_index)
                    object.__setattr__(model, "_paif_pgd_runner", runner)
                delta = runner.attack(vis, ir, label, attack_iters, seed, gidx)
            else:
                delta = pgd_attack_both(model, vis, ir, label, epsilon, alpha, attack_iters, seed, gidx, ignore_index)
            vis, ir = vis + delta.delta_vis, ir + delta.delta_ir
        with torch.no_grad():
            _, seg = model(ir, vis)
            seg = F.interpolate(seg, size=label.shape[1:], mode="bilinear", align_corners=False)
            meter.update(label, seg.argmax(1))
    return meter.all_reduce(group)
