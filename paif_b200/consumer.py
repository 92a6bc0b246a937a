"""Stock-PyTorch stand-ins for what sits either side of the fusion hot path in the reference's task model.

NOT part of the B200 hot path: by the scope contract the segmentation consumer stays stock PyTorch
(SURVEY.md 2 / 8f).  They exist so that the PGD robust-evaluation loop (paif_b200/evaluate.py) can be
run and timed end to end on a GPU box where the reference checkout and its third-party dependencies
(timm, mmcv) do not exist:

* ``SegFormerLite`` — a hierarchical mix-transformer encoder + all-MLP decode head with the MiT-B3
  hyper-parameters the reference selects (``'mit_b3'``, embedding 256, 9 classes; robust_test.py:262,
  core/model_fusion_auto.py:9-68), written from the published SegFormer architecture with
  ``torch.nn.functional.scaled_dot_product_attention``.  Random-init; it reproduces the consumer's
  shape and cost (~100 GFLOP per 480x640 frame), not its trained weights.
* ``FusionSegTask`` — the colour / normalisation glue of ``Network_MM_CompModel.forward``
  (core/model_fusion_auto.py:712-729): RGB->YCrCb, fusion on Y, YCrCb->RGB, clamp, min-max over the
  batch, x255, ImageNet mean/std, consumer.  ``forward(ir, vis) -> (fused, seg_logits)``.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_MEAN = (123.675, 116.28, 103.53)          # core/model_fusion_auto.py:710-711
_STD = (58.395, 57.12, 57.375)


class _GlueFn(torch.autograd.Function):
    """``(fused Y, visible RGB) -> normalised consumer input`` as the two-pass kernels ``paif_glue_forward`` /
    ``paif_glue_backward`` (csrc/glue.cu): Cr/Cb of the visible image, YCrCb -> RGB, clamp, min-max stretch, x255,
    (x - mean) / std (core/model_fusion_auto.py:69-111, 715-727), with the gradients autograd would give."""

    @staticmethod
    def forward(ctx, fused, vis, per_sample):
        B, _, H, W = vis.shape
        fused, vis = fused.contiguous().float(), vis.contiguous().float()
        dev = vis.device
        nblk = _lib.load().paif_glue_blocks(H, W)
        x = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
        partial = torch.empty((B, nblk, 2), device=dev, dtype=torch.float32)
        ties = torch.empty((B, nblk, 2), device=dev, dtype=torch.int32)
        lohi = torch.empty((B, 2), device=dev, dtype=torch.float32)
        mean3, std3 = (ctypes.c_float * 3)(*_MEAN), (ctypes.c_float * 3)(*_STD)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.call("paif_glue_forward", fused.data_ptr(), vis.data_ptr(), mean3, std3, x.data_ptr(), partial.data_ptr(),
                      ties.data_ptr(), lohi.data_ptr(), int(per_sample), B, H, W, stream)
        ctx.save_for_backward(fused, vis, ties, lohi)
        ctx.per_sample, ctx.nblk = int(per_sample), nblk
        return x

    @staticmethod
    def backward(ctx, gx):
        fused, vis, ties, lohi = ctx.saved_tensors
        B, _, H, W = vis.shape
        gx = gx.contiguous().float()
        dev = vis.device
        sums = torch.empty((B, ctx.nblk, 2), device=dev, dtype=torch.float32)
        gfused = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
        gvis = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
        std3 = (ctypes.c_float * 3)(*_STD)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.call("paif_glue_backward", fused.data_ptr(), vis.data_ptr(), gx.data_ptr(), std3, lohi.data_ptr(),
                      ties.data_ptr(), sums.data_ptr(), gfused.data_ptr(), gvis.data_ptr(), ctx.per_sample, B, H, W, stream)
        return gfused, gvis, None


class _MixFFN(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.dw = nn.Conv2d(hidden, hidden, 3, padding=1, groups=hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x, hw):
        h, w = hw
        x = self.fc1(x)
        b, n, c = x.shape
        x = self.dw(x.transpose(1, 2).reshape(b, c, h, w)).flatten(2).transpose(1, 2)
        return self.fc2(F.gelu(x))


class _ReducedAttention(nn.Module):
    """Self-attention whose keys/values come from a spatially reduced copy of the tokens."""

    def __init__(self, dim, heads, reduction):
        super().__init__()
        self.heads = heads
        self.q = nn.Linear(dim, dim)
        self.kv = nn.Linear(dim, 2 * dim)
        self.proj = nn.Linear(dim, dim)
        self.reduction = reduction
        if reduction > 1:
            self.sr = nn.Conv2d(dim, dim, reduction, stride=reduction)
            self.sr_norm = nn.LayerNorm(dim)

    def forward(self, x, hw):
        b, n, c = x.shape
        q = self.q(x).reshape(b, n, self.heads, c // self.heads).transpose(1, 2)
        if self.reduction > 1:
            h, w = hw
            y = self.sr(x.transpose(1, 2).reshape(b, c, h, w)).flatten(2).transpose(1, 2)
            y = self.sr_norm(y)
        else:
            y = x
        kv = self.kv(y).reshape(b, -1, 2, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, kv[0], kv[1])
        return self.proj(o.transpose(1, 2).reshape(b, n, c))


class _Block(nn.Module):
    def __init__(self, dim, heads, reduction, mlp_ratio=4):
        super().__init__()
        self.n1, self.n2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn = _ReducedAttention(dim, heads, reduction)
        self.ffn = _MixFFN(dim, dim * mlp_ratio)

    def forward(self, x, hw):
        x = x + self.attn(self.n1(x), hw)
        return x + self.ffn(self.n2(x), hw)


class SegFormerLite(nn.Module):
    """MiT-B3-shaped encoder (dims 64/128/320/512, depths 3/4/18/3, heads 1/2/5/8, reductions 8/4/2/1)
    + MLP decode head (embedding 256).  ``forward([B,3,H,W]) -> [B,num_classes,H/4,W/4]``."""

    def __init__(self, num_classes=9, embedding_dim=256, dims=(64, 128, 320, 512), depths=(3, 4, 18, 3),
                 heads=(1, 2, 5, 8), reductions=(8, 4, 2, 1)):
        super().__init__()
        self.embeds, self.embed_norms, self.stages, self.stage_norms = (nn.ModuleList() for _ in range(4))
        cin = 3
        for i, d in enumerate(dims):
            k, s = (7, 4) if i == 0 else (3, 2)
            self.embeds.append(nn.Conv2d(cin, d, k, stride=s, padding=k // 2))
            self.embed_norms.append(nn.LayerNorm(d))
            self.stages.append(nn.ModuleList(_Block(d, heads[i], reductions[i]) for _ in range(depths[i])))
            self.stage_norms.append(nn.LayerNorm(d))
            cin = d
        self.head_proj = nn.ModuleList(nn.Linear(d, embedding_dim) for d in dims)
        self.head_fuse = nn.Sequential(nn.Conv2d(4 * embedding_dim, embedding_dim, 1, bias=False),
                                       nn.BatchNorm2d(embedding_dim), nn.ReLU())
        self.head_pred = nn.Conv2d(embedding_dim, num_classes, 1)

    def forward(self, x):
        feats = []
        for embed, enorm, blocks, snorm in zip(self.embeds, self.embed_norms, self.stages, self.stage_norms):
            x = embed(x)
            b, c, h, w = x.shape
            t = enorm(x.flatten(2).transpose(1, 2))
            for blk in blocks:
                t = blk(t, (h, w))
            x = snorm(t).transpose(1, 2).reshape(b, c, h, w)
            feats.append(x)
        h0, w0 = feats[0].shape[2:]
        ups = []
        for f, proj in zip(feats, self.head_proj):
            b, c, h, w = f.shape
            y = proj(f.flatten(2).transpose(1, 2)).transpose(1, 2).reshape(b, -1, h, w)
            ups.append(F.interpolate(y, size=(h0, w0), mode="bilinear", align_corners=False))
        return self.head_pred(self.head_fuse(torch.cat(ups[::-1], 1)))


class FusionSegTask(nn.Module):
    """``Network_MM_CompModel``-style task model (core/model_fusion_auto.py:698-729) around any fusion
    module with ``forward(ir, vis) -> [B,1,H,W]`` and any consumer ``[B,3,H,W] -> logits``."""

    def __init__(self, fusion, consumer, consumer_autocast=None, per_sample_minmax=False, fused_glue=False):
        super().__init__()
        #: True: RGB -> Y inside the fusion net's visible stem (``forward_rgb``) and the whole output-side glue as the
        #: two-pass kernels of csrc/glue.cu (needs the paif_b200 fusion net and CUDA tensors); False: stock PyTorch ops
        self.fused_glue = fused_glue
        self.enhance_net = fusion
        self.denoise_net = consumer
        #: False: min-max over the whole batch, exactly as the reference wrapper (core/model_fusion_auto.py:721-723),
        #: which only ever sees batch 1.  True: min-max per sample = "the reference at B = 1 applied to each sample"
        #: (SURVEY.md 8e caveat 1): what lets an evaluation harness batch frames without coupling them.
        self.per_sample_minmax = per_sample_minmax
        #: optional torch dtype (e.g. torch.bfloat16): run the stock consumer under torch.autocast — one of the
        #: mitigations SURVEY.md 7 allows for the consumer; the fusion path is unaffected (it takes and returns fp32)
        self.consumer_autocast = consumer_autocast
        self.register_buffer("mean", torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1))

    @staticmethod
    def rgb_to_ycrcb(rgb):
        r, g, b = rgb[:, 0:1], rgb[:, 1:2], rgb[:, 2:3]
        y = 0.299 * r + 0.587 * g + 0.114 * b
        return torch.cat([y, (r - y) * 0.713 + 0.5, (b - y) * 0.564 + 0.5], 1)

    @staticmethod
    def ycrcb_to_rgb(ycc):
        y, cr, cb = ycc[:, 0:1], ycc[:, 1:2] - 0.5, ycc[:, 2:3] - 0.5
        r = y + 1.403 * cr
        g = y - 0.714 * cr - 0.344 * cb
        b = y + 1.773 * cb
        return torch.cat([r, g, b], 1)

    def _consume(self, x):
        if self.consumer_autocast is not None:
            with torch.autocast(device_type=x.device.type, dtype=self.consumer_autocast):
                seg = self.denoise_net(x)
            return seg.float()
        return self.denoise_net(x)

    def forward(self, ir, vis):
        if self.fused_glue:
            fused = self.enhance_net.forward_rgb(ir[:, 0:1], vis)
            return fused, self._consume(_GlueFn.apply(fused, vis, self.per_sample_minmax))
        ycc = self.rgb_to_ycrcb(vis)
        fused = self.enhance_net(ir[:, 0:1], ycc[:, 0:1])
        rgb = self.ycrcb_to_rgb(torch.cat([fused, ycc[:, 1:3]], 1)).clamp(0.0, 1.0)
        if self.per_sample_minmax:
            lo, hi = rgb.amin((1, 2, 3), keepdim=True), rgb.amax((1, 2, 3), keepdim=True)
        else:
            lo, hi = rgb.min(), rgb.max()                  # spans the batch, as the reference does (:721-723)
        rgb = (rgb - lo) / (hi - lo).clamp_min(1e-12)
        x = (rgb * 255.0 - self.mean) / self.std
        if self.consumer_autocast is not None:
            with torch.autocast(device_type=x.device.type, dtype=self.consumer_autocast):
                seg = self.denoise_net(x)
            return fused, seg.float()
        return fused, self.denoise_net(x)
