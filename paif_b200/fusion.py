"""Drop-in replacement for PAIF's ``Network_Fusion_Searched``
(reference core/model_fusion_auto.py:599-640) running on hand-written sm_100a kernels
through the C ABI in ``include/paif_b200.h``.

Same constructor ``(C, criterion, genotype_feature, steps=4, multiplier=3)``, same
``state_dict`` keys / shapes (45 for the shipped ``fusion_at`` genotype, so reference
checkpoints load with ``strict=True``), same ``forward(ir, vis) -> [B,1,H,W]`` and
``_loss``; gradients flow to ``ir`` and ``vis`` (backward-to-input only — the PGD loop of
attack/attack.py:417-514 never reads weight gradients).

The nn.Modules below are *parameter containers* that mirror the reference's module tree
(names, construction order and therefore default initialisation); none of their
``nn.Conv2d`` children is ever called.  The arithmetic runs in ``_Runtime`` via ctypes
calls into ``libpaif_b200.so``.  There is no PyTorch/CPU fallback: a missing library or a
CPU tensor raises.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import ConvDesc

_RDB_SCALE = 0.333333       # literal constant, operations_m.py:449


def _ptr(t):
    return None if t is None else t.data_ptr()


def basicconv_padding(k, d):
    """BasicConv padding table, operations_m.py:121-132."""
    return {(3, 1): 1, (3, 2): 2, (5, 1): 2, (5, 2): 4, (7, 1): 3, (7, 2): 6}.get((k, d), 0)


def _check_same_padding(k, d, what):
    if basicconv_padding(k, d) != d * (k - 1) // 2:
        # the reference gives such a (k, d) padding 0, which breaks its own residual adds
        raise NotImplementedError("%s: kernel %d dilation %d is not shape-preserving in the reference" % (what, k, d))


# ----------------------------------------------------------------------------------------
# weight packing (device-side torch ops, re-run only when a parameter changes)
# ----------------------------------------------------------------------------------------
class _ConvW:
    """Packed weights of one dense convolution for both engines."""

    def __init__(self, w, nsrc, k, dil):
        cout, cin_total, kh, kw = w.shape
        assert kh == k and kw == k and cin_total % nsrc == 0
        self.k, self.dil, self.nsrc, self.cout = k, dil, nsrc, cout
        self.cps = cin_total // nsrc
        # direct engine: [nsrc][taps][cin_per_src][cout]
        self.direct = w.reshape(cout, nsrc, self.cps, kh * kw).permute(1, 3, 2, 0).contiguous().float()
        # tcgen05 engine: TF32-rounded UMMA B tiles (layout documented at paif_conv_tc_kq in the header)
        self.mma = None
        kq = _lib.load().paif_conv_tc_kq(nsrc, k, dil) if (cout == 32 and self.cps == 32) else 0
        if kq:
            G = cin_total // (kq * 4)
            # [K-group][dx][k8][16-B chunk][dy][cout][4 cin]: all dy taps of one (dx, k8) form one B tile of N = 32*k rows
            wt = _round_tf32(w.float()).reshape(cout, G, kq // 2, 2, 4, kh, kw)
            self.mma = wt.permute(1, 6, 2, 3, 5, 0, 4).contiguous()
        # bf16 storage mode: the same tile order with 8 bf16 input channels per 16 bytes (paif_conv_tc_kq_bf16)
        self.mma16 = None
        kp = _lib.load().paif_conv_tc_kq_bf16(nsrc, k, dil) if (cout == 32 and self.cps == 32) else 0
        if kp:
            G = cin_total // (kp * 8)
            wt = w.float().reshape(cout, G, kp // 2, 2, 8, kh, kw)
            self.mma16 = wt.permute(1, 6, 2, 3, 5, 0, 4).contiguous().to(torch.bfloat16)


def _round_tf32(w):
    """fp32 -> nearest TF32 (10-bit mantissa), kept in an fp32 container."""
    i = w.contiguous().view(torch.int32)
    return ((i + 0x1000) & -8192).view(torch.float32)


def _dgrad_groups(w, k, dil, scale=None):
    """dgrad of conv(w) as forward convs: one _ConvW per group of 32 input channels.
    gx[ci](q) = sum_{co,t} w[co][ci][flip t] * (scale[co] * gy[co])(q + off_t)."""
    wf = w.flip(-1, -2)
    if scale is not None:
        wf = wf * scale.view(-1, 1, 1, 1)
    cin_total = w.shape[1]
    return [_ConvW(wf[:, g:g + 32].transpose(0, 1).contiguous(), 1, k, dil) for g in range(0, cin_total, 32)]


def _slope(w):
    """PReLU slope as the fp32 scalar the kernels read (a .half() / .double() module still feeds fp32)."""
    return w.detach().float().contiguous()


def _bn_fold(bn):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return s.float().contiguous(), (bn.bias - bn.running_mean * s).float().contiguous()


# ----------------------------------------------------------------------------------------
# runtime: thin wrappers over the C ABI
# ----------------------------------------------------------------------------------------
class _Runtime:
    def __init__(self, B, H, W, C, device, engine, save, bf16=False):
        self.B, self.H, self.W, self.C = B, H, W, C
        self.device = device
        self.engine = engine
        self.save = save
        self.bf16 = bf16             # bf16 C8 activation maps (forward only)
        self.dilconv_dense = True    # DilConv as one dense conv on the tensor-core engine when relu(x) is available
        self.want_relu = False       # set by Cell_Chain: the next op (DilConv) wants relu(out) as a second map ...
        self.last_relu = None        # ... which a producer that can emit it leaves here
        self._zero = None
        self.stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        self.launches = 0
        self.profile = None          # optional list: (name, meta, start_event, end_event) per launch
        self._meta = None

    def zero_slope(self):
        if self._zero is None:
            self._zero = torch.zeros(1, device=self.device, dtype=torch.float32)
        return self._zero

    def tc_engine(self):
        """True when dense convolutions of this call resolve to the tcgen05 engine."""
        return self.engine in (_lib.ENGINE_AUTO, _lib.ENGINE_TCGEN05)

    # buffers ---------------------------------------------------------------------------
    def new_map(self, C=None, fp32=False):
        C = self.C if C is None else C
        if self.bf16 and not fp32:
            return torch.empty((self.B, C // 8, self.H, self.W, 8), device=self.device, dtype=torch.bfloat16)
        return torch.empty((self.B, C // 4, self.H, self.W, 4), device=self.device, dtype=torch.float32)

    def new_plane(self, ch=None):
        shape = (self.B, self.H, self.W) if ch is None else (self.B, self.H, self.W, ch)
        return torch.empty(shape, device=self.device, dtype=torch.float32)

    #: kernels behind one ABI call when it is not exactly one
    _KERNELS_PER_CALL = {"paif_out_forward": 2, "paif_out_forward_bf16": 2, "paif_out_forward_tc": 2, "paif_gf_decomp_backward": 3,
                         "paif_gf_decomp_backward_saved": 3}

    def note_bytes(self, bytes_per_px):
        """algorithmic HBM bytes per pixel of the next launch (inputs read once + outputs written once), for the
        per-kernel roofline table of bench.py"""
        if self.profile is not None:
            self._meta = {"bytes": float(bytes_per_px) * self.B * self.H * self.W}

    def call(self, name, *args):
        self.launches += self._KERNELS_PER_CALL.get(name, 1)
        if self.profile is None:
            _lib.call(name, *args, self.stream)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(name, *args, self.stream)
        e1.record()
        self.profile.append((name, self._meta, e0, e1))
        self._meta = None

    # kernels ---------------------------------------------------------------------------
    def conv(self, srcs, cw, *, ch_scale=None, ch_shift=None, pre_res=(), want_pre=False, mask_src=None,
             mask_slope=None, slope=None, post_scale=1.0, post_res=(), act2_slope=None, want_partials=False,
             src_fp32=False):
        d = ConvDesc()
        d.B, d.H, d.W = self.B, self.H, self.W
        d.nsrc, d.cin_per_src, d.cout = cw.nsrc, cw.cps, cw.cout
        d.kh = d.kw = cw.k
        d.dil = cw.dil
        engine = self.engine
        if engine == _lib.ENGINE_AUTO:
            engine = _lib.ENGINE_TCGEN05 if cw.mma is not None else _lib.ENGINE_DIRECT
        wmma, in_bytes, out_bytes = cw.mma, 4.0, 4.0
        if self.bf16:
            # bf16 maps exist on the tensor-core engine only; src_fp32: fp32 sources (TF32 operands), bf16 everything else
            d.storage = _lib.STORAGE_F32_BF16 if src_fp32 else _lib.STORAGE_BF16
            wmma, in_bytes, out_bytes = (cw.mma, 4.0, 2.0) if src_fp32 else (cw.mma16, 2.0, 2.0)
            if engine != _lib.ENGINE_TCGEN05 or wmma is None:
                raise NotImplementedError("bf16 storage: convolution k=%d dil=%d x%d has no tcgen05 instance" % (cw.k, cw.dil, cw.nsrc))
        d.engine = engine
        assert len(srcs) == cw.nsrc
        for i, s in enumerate(srcs):
            d.src[i] = s.data_ptr()
        d.weight = cw.direct.data_ptr()
        d.weight_mma = _ptr(wmma)
        d.ch_scale, d.ch_shift = _ptr(ch_scale), _ptr(ch_shift)
        assert len(pre_res) <= 2 and len(post_res) <= 3
        for i, r in enumerate(pre_res):
            d.pre_res[i] = r.data_ptr()
        for i, r in enumerate(post_res):
            d.post_res[i] = r.data_ptr()
        out = self.new_map(cw.cout)
        d.out = out.data_ptr()
        pre = act2 = partials = None
        if want_pre:
            pre = self.new_map(cw.cout)
            d.out_pre = pre.data_ptr()
        d.mask_src, d.mask_slope = _ptr(mask_src), _ptr(mask_slope)
        d.slope = _ptr(slope)
        d.post_scale = post_scale
        if act2_slope is not None:
            act2 = self.new_map(cw.cout)
            d.out_act2, d.slope2 = act2.data_ptr(), act2_slope.data_ptr()
        if want_partials:
            tiles = _lib.load().paif_conv_num_tiles(self.H, self.W, engine)
            partials = torch.empty((self.B, tiles, cw.cout), device=self.device, dtype=torch.float32)
            d.chan_partials = partials.data_ptr()
        # algorithmic traffic: every source / residual / mask map read once, every output map written once
        ch_bytes = (in_bytes * cw.nsrc * cw.cps +
                    out_bytes * cw.cout * (len(pre_res) + len(post_res) + (mask_src is not None) + 1 +
                                           bool(want_pre) + (act2_slope is not None)))
        self._meta = {"k": cw.k, "dil": cw.dil, "cin": cw.nsrc * cw.cps, "cout": cw.cout, "engine": engine,
                      "storage": int(d.storage),
                      "flops": 2.0 * cw.cout * cw.nsrc * cw.cps * cw.k * cw.k * self.B * self.H * self.W,
                      "bytes": ch_bytes * self.B * self.H * self.W}
        self.call("paif_conv_forward", ctypes.byref(d))
        return out, pre, act2, partials

    def add(self, a, b, c=None):
        if self.bf16:
            raise NotImplementedError("bf16 storage: this genotype needs a stand-alone map add (fp32 storage only)")
        out = torch.empty_like(a)
        self.call("paif_add_maps", a.data_ptr(), b.data_ptr(), _ptr(c), out.data_ptr(), a.numel())
        return out

    def add_all(self, x, extras):
        extras = list(extras)
        while extras:
            take, extras = extras[:2], extras[2:]
            x = self.add(x, take[0], take[1] if len(take) > 1 else None)
        return x

    def mask_scale(self, g, mask_src, mask_slope, scale):
        out = torch.empty_like(g)
        self.call("paif_mask_scale", g.data_ptr(), _ptr(mask_src), _ptr(mask_slope), scale, out.data_ptr(),
                  g.shape[1] * 4, self.B, self.H, self.W)
        return out

    def dwconv(self, x, w, k, dil, relu_in, mask_src=None, post_res=None):
        out = torch.empty_like(x)
        self.call("paif_dwconv_forward", x.data_ptr(), w.data_ptr(), int(relu_in), _ptr(mask_src), _ptr(post_res),
                  out.data_ptr(), self.C, k, dil, self.B, self.H, self.W)
        return out


# ----------------------------------------------------------------------------------------
# parameter containers mirroring operations_m.py (names and construction order preserved)
# ----------------------------------------------------------------------------------------
class BasicConv(nn.Module):
    """operations_m.py:114-145 (bn / relu sub-modules exist only when requested)."""

    def __init__(self, in_planes, out_planes, kernel_size, dilation=1, groups=1, relu=True, bn=False, bias=False):
        super().__init__()
        self.out_channels = out_planes
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=1,
                              padding=basicconv_padding(kernel_size, dilation), dilation=dilation,
                              groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_planes, eps=1e-5, momentum=0.01, affine=True) if bn else None
        self.relu = nn.PReLU() if relu else None


class _Primitive(nn.Module):
    max_extras = 0       # residual maps the forward epilogue can absorb
    max_extra_add = 0    # gradient maps the backward epilogue can absorb
    bf16_ok = False      # has a bf16-storage forward (the primitives of the shipped genotype do)

    def forward(self, x):
        raise RuntimeError("paif_b200 primitives run only inside Network_Fusion_Searched.forward")


class ResidualDenseBlock(_Primitive):
    """operations_m.py:435-449."""
    max_extras, max_extra_add = 2, 1
    bf16_ok = True

    def __init__(self, in_channels, kernel_size, dialtions=1, bias=False):
        super().__init__()
        _check_same_padding(kernel_size, dialtions, "Denseblocks")
        self.k, self.d = kernel_size, dialtions
        self.conv1 = BasicConv(in_channels, in_channels, kernel_size, dilation=dialtions, relu=False)
        self.conv2 = BasicConv(in_channels * 2, in_channels, kernel_size, dilation=dialtions, relu=False)
        self.conv3 = BasicConv(in_channels * 3, in_channels, kernel_size, dilation=dialtions, relu=False)
        self.lrelu = nn.PReLU()

    def pack(self, need_bwd):
        k, d = self.k, self.d
        p = {"w": [_ConvW(c.conv.weight.detach(), n, k, d)
                   for n, c in ((1, self.conv1), (2, self.conv2), (3, self.conv3))],
             "a": _slope(self.lrelu.weight)}
        if need_bwd:
            p["wd"] = [_dgrad_groups(c.conv.weight.detach(), k, d) for c in (self.conv1, self.conv2, self.conv3)]
        return p

    def fwd(self, rt, p, x, extras):
        a = p["a"]
        x1 = rt.conv([x], p["w"][0], slope=a)[0]
        x2 = rt.conv([x, x1], p["w"][1], slope=a)[0]
        out, pre3, relu_out, _ = rt.conv([x, x1, x2], p["w"][2], slope=a, post_scale=_RDB_SCALE,
                                         post_res=[x] + list(extras), want_pre=rt.save,
                                         act2_slope=rt.zero_slope() if rt.want_relu else None)
        rt.last_relu = relu_out
        return out, (x1, x2, pre3)

    def bwd(self, rt, p, rec, g, extra_add):
        x1, x2, pre3 = rec
        a, wd = p["a"], p["wd"]
        g3 = rt.mask_scale(g, pre3, a, _RDB_SCALE)
        gxa = rt.conv([g3], wd[2][0])[0]
        gx1a = rt.conv([g3], wd[2][1])[0]
        g2 = rt.conv([g3], wd[2][2], mask_src=x2, mask_slope=a)[0]
        gxb = rt.conv([g2], wd[1][0], post_res=[gxa])[0]
        g1 = rt.conv([g2], wd[1][1], pre_res=[gx1a], mask_src=x1, mask_slope=a)[0]
        return rt.conv([g1], wd[0][0], post_res=[gxb, g] + list(extra_add))[0]


class DilConv(_Primitive):
    """operations_m.py:494-506 (BN in eval mode, folded to scale/shift)."""
    max_extras, max_extra_add = 2, 0
    bf16_ok = True

    def __init__(self, C_in, C_out, kernel_size, dilation, affine=True):
        super().__init__()
        _check_same_padding(kernel_size, dilation, "DilConv")
        self.k, self.d = kernel_size, dilation
        self.op = nn.Sequential(
            nn.ReLU(inplace=False),
            BasicConv(C_in, C_out, kernel_size, dilation=dilation, relu=False, groups=C_in),
            nn.Conv2d(C_in, C_out, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(C_out, affine=affine),
        )

    def pack(self, need_bwd):
        dw = self.op[1].conv.weight.detach()
        C = dw.shape[0]
        s, sh = _bn_fold(self.op[3])
        p = {"dw": dw.reshape(C, -1).contiguous().float(), "pw": _ConvW(self.op[2].weight.detach(), 1, 1, 1),
             "pw_raw": self.op[2].weight.detach().reshape(C, C).contiguous().float(), "s": s, "sh": sh}
        # depthwise followed by 1x1 == one dense k x k convolution with rank-1 taps W[co][ci][t] = pw[co][ci] dw[ci][t];
        # on the tensor-core engine that is cheaper than the FFMA kernel whenever the producer can hand over relu(x)
        wd = self.op[2].weight.detach()[:, :, 0, 0].double()[:, :, None, None] * dw.double()[:, 0][None]
        p["wdense"] = _ConvW(wd.float(), 1, self.k, self.d)
        if need_bwd:
            p["dw_t"] = dw.flip(-1, -2).reshape(C, -1).contiguous().float()
            p["pw_d"] = _dgrad_groups(self.op[2].weight.detach(), 1, 1, scale=s)[0]
            p["wdense_d"] = _dgrad_groups(wd.float(), self.k, self.d, scale=s)[0]     # dgrad of the dense form
        return p

    def fwd(self, rt, p, x, extras, x_relu=None):
        extras = list(extras)
        cw = p["wdense"]
        if x_relu is not None and rt.tc_engine() and (cw.mma16 if rt.bf16 else cw.mma) is not None:
            return rt.conv([x_relu], cw, ch_scale=p["s"], ch_shift=p["sh"], post_res=[x] + extras)[0], ()
        if rt.bf16:
            out = torch.empty_like(x)
            rt.call("paif_dilconv_forward_bf16", x.data_ptr(), p["dw"].data_ptr(), p["pw_raw"].data_ptr(),
                    p["s"].data_ptr(), p["sh"].data_ptr(), _ptr(extras[0]) if extras else None,
                    _ptr(extras[1]) if len(extras) > 1 else None, out.data_ptr(), 1, rt.C, self.k, self.d,
                    rt.B, rt.H, rt.W)
            return out, ()
        if (self.k, self.d) not in ((3, 1), (3, 2)):       # only the 3x3 shapes have a fused kernel instance
            t = rt.dwconv(x, p["dw"], self.k, self.d, relu_in=True)
            return rt.conv([t], p["pw"], ch_scale=p["s"], ch_shift=p["sh"], post_res=[x] + extras)[0], ()
        out = torch.empty_like(x)
        rt.call("paif_dilconv_forward", x.data_ptr(), p["dw"].data_ptr(), p["pw_raw"].data_ptr(), p["s"].data_ptr(),
                p["sh"].data_ptr(), _ptr(extras[0]) if extras else None, _ptr(extras[1]) if len(extras) > 1 else None,
                out.data_ptr(), 1, _lib.ENGINE_AUTO, rt.C, self.k, self.d, rt.B, rt.H, rt.W)
        return out, ()

    def bwd(self, rt, p, rec, g, extra_add, x=None):
        cw = p["wdense_d"]
        if rt.tc_engine() and rt.dilconv_dense and cw.mma is not None:
            # one dense dgrad convolution on the engine: ReLU' (mask of x with slope 0) and the "+ x" branch in the epilogue
            return rt.conv([g], cw, mask_src=x, mask_slope=rt.zero_slope(), post_res=[g])[0]
        u = rt.conv([g], p["pw_d"])[0]
        return rt.dwconv(u, p["dw_t"], self.k, self.d, relu_in=False, mask_src=x, post_res=g)


class eca_layer(nn.Module):
    """operations_m.py:340-367 (parameter container)."""

    def __init__(self, channel, c_out, stride, k_size=3):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv = nn.Conv1d(1, 1, kernel_size=k_size, padding=(k_size - 1) // 2, bias=False)
        self.sigmoid = nn.Sigmoid()


class ECABasicBlock(_Primitive):
    """operations_m.py:368-393."""
    max_extras, max_extra_add = 1, 3
    bf16_ok = True

    def __init__(self, inplanes, planes, kernel=3, dilation=1, stride=1, reduction=64, with_norm=False):
        super().__init__()
        _check_same_padding(kernel, 1, "ECAattention")
        self.k = kernel
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.conv2 = BasicConv(inplanes, inplanes, kernel, relu=False)
        self.se = eca_layer(planes, planes, stride, k_size=kernel)
        self.relu = nn.PReLU()

    def pack(self, need_bwd):
        p = {"w1": _ConvW(self.conv1.weight.detach(), 1, 3, 1),
             "w2": _ConvW(self.conv2.conv.weight.detach(), 1, self.k, 1),
             "w1d": self.se.conv.weight.detach().reshape(-1).contiguous().float(),
             "a": _slope(self.relu.weight)}
        if need_bwd:
            p["w1_d"] = _dgrad_groups(self.conv1.weight.detach(), 3, 1)[0]
            p["w2_d"] = _dgrad_groups(self.conv2.conv.weight.detach(), self.k, 1)[0]
        return p

    def fwd(self, rt, p, x, extras):
        a = p["a"]
        x0, _, px0, _ = rt.conv([x], p["w1"], act2_slope=a)
        o, _, _, partials = rt.conv([px0], p["w2"], want_partials=True)
        e = torch.empty((rt.B, rt.C), device=rt.device, dtype=torch.float32)
        rt.call("paif_eca_scale", partials.data_ptr(), partials.shape[1], p["w1d"].data_ptr(), self.k,
                e.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        out = rt.new_map()
        rt.note_bytes((2 if rt.bf16 else 4) * rt.C * (3 + len(extras[:1])))
        rt.call("paif_eca_apply_bf16" if rt.bf16 else "paif_eca_apply", o.data_ptr(), x0.data_ptr(), e.data_ptr(),
                a.data_ptr(), _ptr(extras[0]) if extras else None, out.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        return out, (x0, o, e)

    def bwd(self, rt, p, rec, g, extra_add):
        x0, o, e = rec
        a = p["a"]
        lib = _lib.load()
        tiles = lib.paif_eca_bwd_tiles(rt.H, rt.W)
        gw = rt.new_map()
        partials = torch.empty((rt.B, tiles, rt.C), device=rt.device, dtype=torch.float32)
        rt.call("paif_eca_bwd_pass1", g.data_ptr(), o.data_ptr(), x0.data_ptr(), e.data_ptr(), a.data_ptr(),
                gw.data_ptr(), partials.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        gm = torch.empty((rt.B, rt.C), device=rt.device, dtype=torch.float32)
        rt.call("paif_eca_bwd_scale", partials.data_ptr(), tiles, e.data_ptr(), p["w1d"].data_ptr(), self.k,
                gm.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        go = rt.new_map()
        rt.call("paif_eca_bwd_pass2", gw.data_ptr(), e.data_ptr(), gm.data_ptr(), go.data_ptr(),
                rt.C, rt.B, rt.H, rt.W)
        gx0 = rt.conv([go], p["w2_d"], mask_src=x0, mask_slope=a, post_res=[gw])[0]
        return rt.conv([gx0], p["w1_d"], post_res=list(extra_add))[0]


class ResidualModule(_Primitive):
    """operations_m.py:451-464.  The bias-free 3x3(dil 2) and 1x1 convolutions that follow the
    k x k convolution have no activation between them and are merged into one 3x3(dil 2)."""
    max_extras, max_extra_add = 2, 2
    bf16_ok = True

    def __init__(self, in_channels, kernel_size, dialtions=1, bias=False):
        super().__init__()
        _check_same_padding(kernel_size, dialtions, "Residualblocks")
        self.k, self.d = kernel_size, dialtions
        self.op = nn.Sequential(
            BasicConv(in_channels, in_channels, kernel_size, dilation=dialtions, relu=False),
            nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=2, dilation=2, bias=False),
            nn.Conv2d(in_channels, in_channels, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(in_channels),
            nn.PReLU(),
        )

    def pack(self, need_bwd):
        w0 = self.op[0].conv.weight.detach()
        wm = torch.einsum('om,mikl->oikl', self.op[2].weight.detach()[:, :, 0, 0].double(),
                          self.op[1].weight.detach().double()).float()
        s, sh = _bn_fold(self.op[3])
        p = {"w0": _ConvW(w0, 1, self.k, self.d), "wm": _ConvW(wm, 1, 3, 2), "s": s, "sh": sh,
             "a": _slope(self.op[4].weight)}
        if need_bwd:
            p["w0_d"] = _dgrad_groups(w0, self.k, self.d)[0]
            p["wm_d"] = _dgrad_groups(wm, 3, 2, scale=s)[0]
        return p

    def fwd(self, rt, p, x, extras):
        t1 = rt.conv([x], p["w0"])[0]
        out, pre, _, _ = rt.conv([t1], p["wm"], ch_scale=p["s"], ch_shift=p["sh"], slope=p["a"],
                                 post_res=[x] + list(extras), want_pre=rt.save)
        return out, (pre,)

    def bwd(self, rt, p, rec, g, extra_add, g_masked=None):
        (pre,) = rec
        gm = g_masked if g_masked is not None else rt.mask_scale(g, pre, p["a"], 1.0)
        gt1 = rt.conv([gm], p["wm_d"])[0]
        return rt.conv([gt1], p["w0_d"], post_res=[g] + list(extra_add))[0]


class SepConv(_Primitive):
    """operations_m.py:509-526: two (ReLU -> depthwise k x k -> 1x1 -> BN) halves, no residual of its own.
    OPS passes padding k//2 and drops the dilation (operations_m.py:15)."""
    max_extras, max_extra_add = 2, 0

    def __init__(self, C_in, C_out, kernel_size, stride, padding, affine=True):
        super().__init__()
        if stride != 1 or padding != kernel_size // 2 or kernel_size % 2 == 0 or kernel_size > 7:
            raise NotImplementedError("SepConv: only odd k <= 7, stride 1, 'same' padding")
        self.k = kernel_size
        self.op = nn.Sequential(
            nn.ReLU(inplace=False),
            nn.Conv2d(C_in, C_in, kernel_size=kernel_size, stride=stride, padding=padding, groups=C_in, bias=False),
            nn.Conv2d(C_in, C_in, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(C_in, affine=affine),
            nn.ReLU(inplace=False),
            nn.Conv2d(C_in, C_in, kernel_size=kernel_size, stride=1, padding=padding, groups=C_in, bias=False),
            nn.Conv2d(C_in, C_out, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(C_out, affine=affine),
        )

    def pack(self, need_bwd):
        halves = []
        for dw_m, pw_m, bn_m in ((self.op[1], self.op[2], self.op[3]), (self.op[5], self.op[6], self.op[7])):
            dw = dw_m.weight.detach()
            C = dw.shape[0]
            s, sh = _bn_fold(bn_m)
            h = {"dw": dw.reshape(C, -1).contiguous().float(), "pw": _ConvW(pw_m.weight.detach(), 1, 1, 1),
                 "pw_raw": pw_m.weight.detach().reshape(C, C).contiguous().float(), "s": s, "sh": sh}
            if need_bwd:
                h["dw_t"] = dw.flip(-1, -2).reshape(C, -1).contiguous().float()
                h["pw_d"] = _dgrad_groups(pw_m.weight.detach(), 1, 1, scale=s)[0]
            halves.append(h)
        return {"halves": halves}

    def _half(self, rt, h, x, extras):
        extras = list(extras)
        if self.k != 3:                                     # only 3x3 has a fused kernel instance
            t = rt.dwconv(x, h["dw"], self.k, 1, relu_in=True)
            return rt.conv([t], h["pw"], ch_scale=h["s"], ch_shift=h["sh"], post_res=extras)[0]
        out = torch.empty_like(x)
        rt.call("paif_dilconv_forward", x.data_ptr(), h["dw"].data_ptr(), h["pw_raw"].data_ptr(), h["s"].data_ptr(),
                h["sh"].data_ptr(), _ptr(extras[0]) if extras else None, _ptr(extras[1]) if len(extras) > 1 else None,
                out.data_ptr(), 0, _lib.ENGINE_AUTO, rt.C, self.k, 1, rt.B, rt.H, rt.W)
        return out

    def fwd(self, rt, p, x, extras):
        y1 = self._half(rt, p["halves"][0], x, [])
        return self._half(rt, p["halves"][1], y1, extras), (y1,)

    def bwd(self, rt, p, rec, g, extra_add, x=None):
        (y1,) = rec
        h1, h2 = p["halves"]
        u = rt.conv([g], h2["pw_d"])[0]
        gy1 = rt.dwconv(u, h2["dw_t"], self.k, 1, relu_in=False, mask_src=y1)
        u = rt.conv([gy1], h1["pw_d"])[0]
        return rt.dwconv(u, h1["dw_t"], self.k, 1, relu_in=False, mask_src=x)


class spatial_attn_layer(nn.Module):
    """operations_m.py:152-163 (1-arg ChannelPool + BasicConv(2, 1, k) + sigmoid); parameter container."""

    def __init__(self, kernel_size=5):
        super().__init__()
        self.compress = nn.Module()
        self.spatial = BasicConv(2, 1, kernel_size, relu=False)


class Spatial_BasicBlock(_Primitive):
    """operations_m.py:179-206 (with_norm=False).  The 1-arg spatial attention out * sigma(conv([max_c, mean_c]))
    runs on the 2-map blend kernels with an all-zero second map and zero weights for its pooled channels."""
    max_extras, max_extra_add = 0, 3

    def __init__(self, inplanes, planes, kernel=3, dilation=1, stride=1, reduction=64, with_norm=False):
        super().__init__()
        _check_same_padding(kernel, 1, "SPAattention")
        self.k = kernel
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.conv2 = BasicConv(inplanes, inplanes, kernel, relu=False)
        self.se = spatial_attn_layer(kernel)
        self.relu = nn.PReLU()

    def pack(self, need_bwd):
        w = self.se.spatial.conv.weight.detach()
        w4 = torch.cat([w.reshape(2, -1), torch.zeros_like(w.reshape(2, -1))], 0).contiguous().float()
        p = {"w1": _ConvW(self.conv1.weight.detach(), 1, 3, 1),
             "w2": _ConvW(self.conv2.conv.weight.detach(), 1, self.k, 1),
             "w4": w4, "a": _slope(self.relu.weight)}
        if need_bwd:
            p["w1_d"] = _dgrad_groups(self.conv1.weight.detach(), 3, 1)[0]
            p["w2_d"] = _dgrad_groups(self.conv2.conv.weight.detach(), self.k, 1)[0]
        return p

    def fwd(self, rt, p, x, extras):
        a = p["a"]
        x0, _, px0, _ = rt.conv([x], p["w1"], act2_slope=a)
        o = rt.conv([px0], p["w2"])[0]
        zero = torch.zeros_like(o)
        pooled = rt.new_plane(4)
        rt.call("paif_channel_pool", o.data_ptr(), zero.data_ptr(), pooled.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        os_, scale = rt.new_map(), rt.new_plane()
        rt.call("paif_spa_blend_forward", pooled.data_ptr(), p["w4"].data_ptr(), self.k, o.data_ptr(), zero.data_ptr(),
                os_.data_ptr(), scale.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        out = rt.new_map()
        pre = rt.new_map() if rt.save else None
        rt.call("paif_add_act", os_.data_ptr(), x0.data_ptr(), a.data_ptr(), out.data_ptr(), _ptr(pre), out.numel())
        return out, (x0, o, scale, pre)

    def bwd(self, rt, p, rec, g, extra_add):
        x0, o, scale, pre = rec
        a = p["a"]
        gpre = rt.mask_scale(g, pre, a, 1.0)
        zero = torch.zeros_like(o)
        gplane = rt.new_plane()
        rt.call("paif_spa_blend_backward_pre", gpre.data_ptr(), o.data_ptr(), zero.data_ptr(), scale.data_ptr(),
                gplane.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        g_o, g_dummy = rt.new_map(), rt.new_map()
        rt.call("paif_spa_blend_backward", gpre.data_ptr(), o.data_ptr(), zero.data_ptr(), scale.data_ptr(),
                gplane.data_ptr(), p["w4"].data_ptr(), self.k, g_o.data_ptr(), g_dummy.data_ptr(), rt.C, rt.B, rt.H, rt.W)
        gx0 = rt.conv([g_o], p["w2_d"], mask_src=x0, mask_slope=a, post_res=[gpre])[0]
        return rt.conv([gx0], p["w1_d"], post_res=list(extra_add))[0]


def _not_on_path(name):
    def make(*a, **k):
        raise NotImplementedError(
            "primitive %r is in the reference's OPS table (operations_m.py:9-18) but not implemented by "
            "paif_b200 (only the shipped fusion_at genotype's primitives are)" % name)
    return make


# OPS, operations_m.py:9-18 (the `affine` flag is dropped there too)
OPS = {
    'Denseblocks': lambda C, kernel, dialtion, affine: ResidualDenseBlock(C, kernel, dialtion),
    'Residualblocks': lambda C, kernel, dialtion, affine: ResidualModule(C, kernel, dialtion),
    'ECAattention': lambda C, kernel, dialtion, affine: ECABasicBlock(C, C, kernel, dialtion),
    'SPAattention': lambda C, kernel, dialtion, affine: Spatial_BasicBlock(C, C, kernel, dialtion),
    'DilConv': lambda C, kernel, dialtion, affine: DilConv(C, C, kernel, dialtion),
    'SepConv': lambda C, kernel, dialtion, affine: SepConv(C, C, kernel, 1, kernel // 2),
    'SelAttention': _not_on_path('SelAttention'),
}


class MixedOp(nn.Module):
    """core/model_fusion_auto.py:397-415 (name grammar ``Name_k[_d]``)."""

    def __init__(self, C, primitive):
        super().__init__()
        self._ops = nn.ModuleList()
        kernel, dilation = 3, 1
        parts = primitive.split('_')
        name, kernel = parts[0], int(parts[1])
        if primitive.find('attention') == -1:
            dilation = int(parts[2])
        self.primitive = primitive
        self._op = OPS[name](C, kernel, dilation, False)


class Cell_Chain(nn.Module):
    """core/model_fusion_auto.py:418-445: out = inp + op_k(...op_1(inp)) (indices/concat unused)."""

    def __init__(self, C, type, concat):
        super().__init__()
        op_names, indices = zip(*type)
        assert len(op_names) == len(indices)
        self._steps = len(op_names)
        self._concat = concat
        self.multiplier = len(concat)
        self._ops = nn.ModuleList()
        for name in op_names:
            self._ops += [MixedOp(C, name)]
        self._indices = indices

    def pack(self, need_bwd):
        return [m._op.pack(need_bwd) for m in self._ops]

    def fwd(self, rt, packs, x, extras):
        """returns (inp + ops(inp) + sum(extras), records)."""
        s, recs = x, []
        n = len(self._ops)
        rt.last_relu = None
        for i, m in enumerate(self._ops):
            op = m._op
            want = ([x] + list(extras)) if i == n - 1 else []
            fused, rest = want[:op.max_extras], want[op.max_extras:]
            inp = s
            relu_in, rt.last_relu = rt.last_relu, None
            nxt = self._ops[i + 1]._op if i + 1 < n else None
            rt.want_relu = isinstance(nxt, DilConv) and rt.tc_engine() and rt.dilconv_dense
            if isinstance(op, DilConv):
                s, rec = op.fwd(rt, packs[i], s, fused, x_relu=relu_in)
            else:
                s, rec = op.fwd(rt, packs[i], s, fused)
            rt.want_relu = False
            if rest:
                s = rt.add_all(s, rest)
                rt.last_relu = None
            recs.append((inp if rt.save else None, rec))
        return s, recs

    def bwd(self, rt, packs, recs, g):
        """gradient w.r.t. the chain input (the chain residual included); extras' gradient is g."""
        gs = g
        for i in reversed(range(len(self._ops))):
            op = self._ops[i]._op
            inp, rec = recs[i]
            want = [g] if i == 0 else []
            fused, rest = want[:op.max_extra_add], want[op.max_extra_add:]
            if isinstance(op, (DilConv, SepConv)):
                gs = op.bwd(rt, packs[i], rec, gs, fused, x=inp)
            else:
                gs = op.bwd(rt, packs[i], rec, gs, fused)
            if rest:
                gs = rt.add_all(gs, rest)
        return gs


class ChannelPool(nn.Module):
    """2-arg ChannelPool, core/model_fusion_auto.py:1352-1355 (fused into the runtime)."""


class spatial_attn_layer_M(nn.Module):
    """core/model_fusion_auto.py:1358-1368."""

    def __init__(self, kernel_size=5):
        super().__init__()
        self.compress = ChannelPool()
        self.spatial = BasicConv(4, 1, kernel_size, relu=False)


class Cell_Decom(nn.Module):
    """core/model_fusion_auto.py:492-535."""

    def __init__(self, C, types, concat):
        super().__init__()
        self._C = C
        self.radiux = [4]
        self.eps_list = [0.001, 0.0001]
        self._ops_1 = nn.ModuleList()
        self._ops_2 = nn.ModuleList()
        self.conv1x1_lf = nn.Conv2d(C * 4, C, kernel_size=1, bias=True)
        self.conv1x1_hf = nn.Conv2d(C * 4, C, kernel_size=1, bias=True)
        self._steps = len(concat)
        self.relu = nn.PReLU()      # present in the state_dict, unused by forward (reference quirk)
        self.chain = Cell_Chain(C, types[0], concat)
        self.chain2 = Cell_Chain(C, types[1], concat)


def _fold_decomp_1x1(w, double=False):
    """conv1x1(cat[LF1, LF2, z-LF1, z-LF2]) == conv1x1'(cat[LF1, LF2, z]) (core/model_fusion_auto.py:512, :533-534)."""
    C = w.shape[0]
    w = w.detach()[:, :, 0, 0].double()
    wa = w[:, 0:C] - w[:, 2 * C:3 * C]
    wb = w[:, C:2 * C] - w[:, 3 * C:4 * C]
    wc = w[:, 2 * C:3 * C] + w[:, 3 * C:4 * C]
    if double:
        return wa, wb, wc
    return torch.cat([wa, wb, wc], 1).float().reshape(C, 3 * C, 1, 1)


def _pack_gf_mix(w):
    """Weight image of ``paif_gf_mix_forward``: TF32-rounded UMMA B tiles ``[K8 step][16-B chunk][n][4 k]`` of
    [Wa; Wb] (n = 64), Wa + Wb (n = 32) and Wc (n = 32), 16 KB (see include/paif_b200.h)."""
    wa, wb, wc = _fold_decomp_1x1(w, double=True)

    def tile(m):                                     # [n][32 k] -> [4 steps][2 chunks][n][4]
        return _round_tf32(m.float()).reshape(m.shape[0], 4, 2, 4).permute(1, 2, 0, 3).contiguous().reshape(-1)

    return torch.cat([tile(torch.cat([wa, wb], 0)), tile(wa + wb), tile(wc)]).contiguous()


def _merge_stem_out(w1, w2):
    """Merge Conv(C->C/2,3x3,pad1) and Conv(C/2->1,3x3,pad1) (core/model_fusion_auto.py:615-618) into 5x5
    stencils wm[3][3][25][C]; class (cy, cx) excludes the second conv's taps that would read its
    zero padding: cy==0 drops dy2=-1, cy==2 drops dy2=+1 (same for x)."""
    w1 = w1.detach().double()          # [M][C][3][3]
    w2 = w2.detach().double()[0]       # [M][3][3]
    C = w1.shape[1]
    wm = torch.zeros(3, 3, 5, 5, C, dtype=torch.float64, device=w1.device)
    for cy in range(3):
        for cx in range(3):
            for ty2 in range(3):
                if (cy == 0 and ty2 == 0) or (cy == 2 and ty2 == 2):
                    continue
                for tx2 in range(3):
                    if (cx == 0 and tx2 == 0) or (cx == 2 and tx2 == 2):
                        continue
                    # tap t = t1 + t2: rows ty2..ty2+2, cols tx2..tx2+2 of the 5x5 window
                    contrib = torch.einsum('m,mikl->kli', w2[:, ty2, tx2], w1)
                    wm[cy, cx, ty2:ty2 + 3, tx2:tx2 + 3] += contrib
    return wm.reshape(3, 3, 25, C).float().contiguous()


class Network_Fusion_Searched(nn.Module):
    """Drop-in for core/model_fusion_auto.py:599-640."""

    def __init__(self, C, criterion, genotype_feature, steps=4, multiplier=3):
        super().__init__()
        if C != 32:
            raise NotImplementedError("paif_b200 kernels are built for C = 32 (the only width the reference ships)")
        self._C = C
        self._criterion = criterion
        self._steps = steps
        self._multiplier = multiplier
        self._genotype = genotype_feature
        self.stem_1 = nn.Sequential(nn.Conv2d(1, C, 3, padding=1, bias=False), nn.PReLU())
        self.stem_2 = nn.Sequential(nn.Conv2d(1, C, 3, padding=1, bias=False), nn.PReLU())
        self.stem_out = nn.Sequential(
            nn.Conv2d(C, C // 2, 3, padding=1, bias=False),
            nn.Conv2d(C // 2, 1, 3, padding=1, bias=False),
            nn.PReLU(),
        )
        self.tanh = nn.Tanh()
        self.spa = spatial_attn_layer_M()
        self.decompation = Cell_Decom(C, [self._genotype.normal_1, self._genotype.normal_2],
                                      self._genotype.normal_1_concat)
        self.chain = Cell_Chain(C, self._genotype.normal_3, self._genotype.normal_1_concat)
        #: 'auto' | 'direct' (exact fp32 FFMA) | 'tcgen05' (TF32 tensor cores, fp32 accumulate)
        self.conv_engine = 'auto'
        #: activation storage: 'fp32' (default; TF32 tensor-core operands, max-abs 1e-3 tier) or 'bf16' (bf16 C8 maps
        #: and bf16 operands after the decomposition, fp32 accumulation; forward-only, north_star's 1e-2 tier)
        self.storage = 'fp32'
        #: run DilConv as one dense k x k convolution on the tensor-core engine (needs relu(x) from the producing
        #: op's epilogue; False keeps the FFMA depthwise+1x1 kernel, which is also what conv_engine='direct' uses)
        self.dilconv_dense = True
        #: stem_out's merged 5x5 stencil on the tensor-core engine (False / conv_engine='direct': the FFMA kernel)
        self.out_tensor_core = True
        #: inference (no saved activations) of the shipped genotype as ONE C-ABI call, ``paif_fusion_forward`` — the same
        #: kernels in the same order as the per-operator path, bit-identical results, no Python between the launches
        self.native_forward = True
        #: decomposition + folded 1x1 in one kernel (paif_gf_mix_forward: channel mix between the two box-filter levels
        #: on tcgen05, no LF maps); False / conv_engine='direct' / forward2: guided filter, then the 1x1 as a convolution
        self.gf_fused = os.environ.get("PAIF_GF_FUSED", "1") != "0"      # (the env switch exists for compute-sanitizer runs)
        self._pack_cache = None
        self._pack_epoch = 0
        self.last_launches = 0
        #: set to a list to collect (name, meta, start_event, end_event) for every kernel launch
        self.profile = None

    # -------------------------------------------------------------------------------------
    def _pack_key(self, need_bwd):
        ts = list(self.parameters()) + list(self.buffers())
        return (need_bwd,) + tuple((t.data_ptr(), t._version) for t in ts)

    def invalidate_packed(self):
        """Drop the packed / TF32-rounded / BN-folded weight cache.  The cache key is (data_ptr, _version) per
        parameter and buffer, which in-place edits through ``.data`` (``p.data.copy_()``, older EMA / checkpoint
        code) do not bump: call this after such an edit.  ``load_state_dict`` and ``.to()/.cuda()/.half()`` clear
        the cache by themselves."""
        self._pack_cache = None
        self._pack_epoch = getattr(self, "_pack_epoch", 0) + 1

    def pack_signature(self, need_bwd=True):
        """Hashable identity of the packed weights a forward would use now (parameter versions, engine switches):
        holders of device pointers into the pack (CUDA-graph replays) compare it before reuse."""
        return (self._pack_key(need_bwd), getattr(self, "_pack_epoch", 0), self.conv_engine, self.storage,
                self.dilconv_dense, self.out_tensor_core, self.gf_fused, self.native_forward)

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def _packed(self, need_bwd):
        key = self._pack_key(need_bwd)
        if self._pack_cache is not None and self._pack_cache[0] == key:
            return self._pack_cache[1]
        if self._pack_cache is not None and self._pack_cache[0][1:] == key[1:] and self._pack_cache[0][0]:
            return self._pack_cache[1]      # a backward-capable pack also serves forward-only calls
        with torch.no_grad():
            d = self.decompation
            p = {
                "stem_w": [s[0].weight.detach().reshape(self._C, 9).contiguous().float() for s in (self.stem_1, self.stem_2)],
                "stem_a": [_slope(s[1].weight) for s in (self.stem_1, self.stem_2)],
                "c1x1": [_ConvW(_fold_decomp_1x1(c.weight), 3, 1, 1) for c in (d.conv1x1_lf, d.conv1x1_hf)],
                "c1x1_b": [c.bias.detach().float().contiguous() for c in (d.conv1x1_lf, d.conv1x1_hf)],
                "gfmix_w": [_pack_gf_mix(c.weight) for c in (d.conv1x1_lf, d.conv1x1_hf)],
                "chain_ir": d.chain.pack(need_bwd), "chain_vis": d.chain2.pack(need_bwd),
                "chain": self.chain.pack(need_bwd),
                "spa_w": self.spa.spatial.conv.weight.detach().reshape(4, -1).contiguous().float(),
                "spa_k": self.spa.spatial.conv.weight.shape[-1],
                "out_wm": _merge_stem_out(self.stem_out[0].weight, self.stem_out[1].weight),
                "out_a": _slope(self.stem_out[2].weight),
            }
            # interior class (1, 1) of the merged stencil as a tensor-core weight image, cout padded 1 -> 16
            # (paif_out_forward_tc); same tile order as _ConvW.mma / .mma16
            C_ = self._C
            wo = torch.zeros(16, C_, 5, 5, device=p["out_wm"].device)
            wo[0] = p["out_wm"][1, 1].reshape(5, 5, C_).permute(2, 0, 1)
            p["out_mma"] = _round_tf32(wo).reshape(16, 1, C_ // 8, 2, 4, 5, 5).permute(1, 6, 2, 3, 5, 0, 4).contiguous()
            p["out_mma16"] = wo.reshape(16, 1, C_ // 16, 2, 8, 5, 5).permute(1, 6, 2, 3, 5, 0, 4).contiguous().to(torch.bfloat16)
            if need_bwd:
                p["c1x1_d"] = [_dgrad_groups(_fold_decomp_1x1(c.weight), 1, 1) for c in (d.conv1x1_lf, d.conv1x1_hf)]
                slopes = [t for n, t in self.named_parameters() if t.numel() == 1 and n != 'decompation.relu.weight']
                if slopes and not bool((torch.cat([t.detach().reshape(1) for t in slopes]) > 0).all()):
                    raise NotImplementedError("backward-to-input needs every PReLU slope > 0 "
                                              "(masks are rebuilt from saved activations)")
        self._pack_cache = (key, p)
        return p

    # -------------------------------------------------------------------------------------
    def _native_weights(self, p):
        """``PaifFusionWeights`` over the packed tensors of ``p`` (kept alive by the pack cache), or None when this
        module is not the shipped ``fusion_at`` structure that ``paif_fusion_forward`` implements."""
        if "native" in p:
            return p["native"]
        d = self.decompation
        ok = ([type(m._op) for m in d.chain._ops] == [ResidualDenseBlock, DilConv]
              and [type(m._op) for m in d.chain2._ops] == [ResidualDenseBlock, ResidualDenseBlock]
              and [type(m._op) for m in self.chain._ops] == [ECABasicBlock, ResidualModule]
              and all((o.k, o.d) == (3, 1) for o in (d.chain._ops[0]._op, d.chain2._ops[0]._op, d.chain2._ops[1]._op))
              and (d.chain._ops[1]._op.k, d.chain._ops[1]._op.d) == (3, 2)
              and self.chain._ops[0]._op.k == 3 and (self.chain._ops[1]._op.k, self.chain._ops[1]._op.d) == (7, 1)
              and p["spa_k"] == 5)
        w = None
        if ok:
            w = _lib.FusionWeights()

            def cv(dst, cw):
                dst.direct, dst.mma_tf32, dst.mma_bf16 = _ptr(cw.direct), _ptr(cw.mma), _ptr(cw.mma16)

            for i in range(2):
                w.stem_w[i], w.stem_a[i] = p["stem_w"][i].data_ptr(), p["stem_a"][i].data_ptr()
                w.gfmix_w[i], w.c1x1_b[i] = p["gfmix_w"][i].data_ptr(), p["c1x1_b"][i].data_ptr()
            for i, pk in enumerate((p["chain_ir"][0], p["chain_vis"][0], p["chain_vis"][1])):
                for j in range(3):
                    cv(w.rdb[i].conv[j], pk["w"][j])
                w.rdb[i].slope = pk["a"].data_ptr()
            dil = p["chain_ir"][1]
            cv(w.dil_dense, dil["wdense"])
            w.dil_scale, w.dil_shift = dil["s"].data_ptr(), dil["sh"].data_ptr()
            w.spa_w, w.spa_k = p["spa_w"].data_ptr(), p["spa_k"]
            eca, res = p["chain"]
            cv(w.eca_conv1, eca["w1"])
            cv(w.eca_conv2, eca["w2"])
            w.eca_w1d, w.eca_a = eca["w1d"].data_ptr(), eca["a"].data_ptr()
            cv(w.res_conv7, res["w0"])
            cv(w.res_merged, res["wm"])
            w.res_scale, w.res_shift, w.res_a = res["s"].data_ptr(), res["sh"].data_ptr(), res["a"].data_ptr()
            w.out_mma_tf32, w.out_mma_bf16 = p["out_mma"].data_ptr(), p["out_mma16"].data_ptr()
            w.out_wm, w.out_a = p["out_wm"].data_ptr(), p["out_a"].data_ptr()
        p["native"] = w
        return w

    def _native_grad_weights(self, p):
        """``PaifFusionGradWeights`` over the dgrad packs of ``p`` (a backward-capable pack of the shipped genotype)."""
        if "native_grad" in p:
            return p["native_grad"]
        g = _lib.FusionGradWeights()

        def cv(dst, cw):
            dst.direct, dst.mma_tf32, dst.mma_bf16 = _ptr(cw.direct), _ptr(cw.mma), _ptr(cw.mma16)

        for i, pk in enumerate((p["chain_ir"][0], p["chain_vis"][0], p["chain_vis"][1])):
            wd = pk["wd"]
            for j in range(3):
                cv(g.rdb[i].c3[j], wd[2][j])
            for j in range(2):
                cv(g.rdb[i].c2[j], wd[1][j])
            cv(g.rdb[i].c1, wd[0][0])
        cv(g.dil_dense_d, p["chain_ir"][1]["wdense_d"])
        eca, res = p["chain"]
        cv(g.eca_conv1_d, eca["w1_d"])
        cv(g.eca_conv2_d, eca["w2_d"])
        cv(g.res_conv7_d, res["w0_d"])
        cv(g.res_merged_d, res["wm_d"])
        for i in range(2):
            for j in range(3):
                cv(g.c1x1_d[i][j], p["c1x1_d"][i][j])
        p["native_grad"] = g
        return g

    def _run_forward_native_save(self, ir, vis):
        """Forward that keeps its activations for ``paif_fusion_backward_input`` as ONE C-ABI call
        (``paif_fusion_forward_save``) over one workspace tensor; None when the configuration is outside what the
        whole-network entry points implement (the per-operator path with saved tensors runs instead)."""
        if not (self.native_forward and self.conv_engine in ('auto', 'tcgen05') and self.gf_fused and self.dilconv_dense
                and self.out_tensor_core and self.profile is None and self.storage == 'fp32'):
            return None
        B, _, H, W = ir.shape
        lib = _lib.load()
        if not lib.paif_gf_mix_supported(self._C, H, W):
            return None
        p = self._packed(True)
        w = self._native_weights(p)
        if w is None:
            return None
        gw = self._native_grad_weights(p)
        need = lib.paif_fusion_train_workspace_bytes(B, H, W)
        ws = torch.empty((need,), device=ir.device, dtype=torch.uint8)
        out = torch.empty((B, 1, H, W), device=ir.device, dtype=torch.float32)
        stream = ctypes.c_void_p(torch.cuda.current_stream(ir.device).cuda_stream)
        _lib.call("paif_fusion_forward_save", ctypes.byref(w), ir.data_ptr(), ir.stride(0), ir.stride(2), ir.stride(3),
                  vis.data_ptr(), vis.stride(0), vis.stride(2), vis.stride(3), out.data_ptr(), ws.data_ptr(), need,
                  B, H, W, stream)
        self.last_launches = 28
        return out, dict(B=B, H=H, W=W, native_ws=ws, native_w=w, native_gw=gw, packed=p, out=out)

    def _run_backward_native(self, saved, g):
        B, H, W = saved["B"], saved["H"], saved["W"]
        ws = saved["native_ws"]
        g_ir = torch.empty((B, H, W), device=g.device, dtype=torch.float32)
        g_vis = torch.empty((B, H, W), device=g.device, dtype=torch.float32)
        stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        _lib.call("paif_fusion_backward_input", ctypes.byref(saved["native_w"]), ctypes.byref(saved["native_gw"]),
                  g.data_ptr(), g_ir.data_ptr(), g_vis.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W, stream)
        self.last_launches = 28 + 47
        return g_ir, g_vis

    def _run_forward_native(self, ir, vis):
        """The whole forward as ONE C-ABI call (``paif_fusion_forward``) over one workspace tensor; None when the
        configuration is outside what that entry point implements (the per-operator path runs instead)."""
        if not (self.native_forward and self.conv_engine in ('auto', 'tcgen05') and self.gf_fused and self.dilconv_dense
                and self.out_tensor_core and self.profile is None):
            return None
        B, _, H, W = ir.shape
        lib = _lib.load()
        if not lib.paif_gf_mix_supported(self._C, H, W):
            return None
        p = self._packed(False)
        w = self._native_weights(p)
        if w is None:
            return None
        storage = _lib.STORAGE_BF16 if self._bf16_storage(False) else _lib.STORAGE_F32
        need = lib.paif_fusion_workspace_bytes(B, H, W, storage)
        ws = torch.empty((need,), device=ir.device, dtype=torch.uint8)
        out = torch.empty((B, 1, H, W), device=ir.device, dtype=torch.float32)
        stream = ctypes.c_void_p(torch.cuda.current_stream(ir.device).cuda_stream)
        _lib.call("paif_fusion_forward", ctypes.byref(w), ir.data_ptr(), ir.stride(0), ir.stride(2), ir.stride(3),
                  vis.data_ptr(), vis.stride(0), vis.stride(2), vis.stride(3), out.data_ptr(), ws.data_ptr(), need,
                  storage, B, H, W, stream)
        self.last_launches = 28
        return out

    def _bf16_storage(self, save):
        """True when this forward runs with bf16 activation maps (``self.storage == 'bf16'``)."""
        if self.storage == 'fp32':
            return False
        if self.storage != 'bf16':
            raise ValueError("storage must be 'fp32' or 'bf16', not %r" % (self.storage,))
        if self.conv_engine == 'direct':
            raise RuntimeError("storage='bf16' runs on the tcgen05 engine; conv_engine='direct' is the exact-fp32 path")
        if self._C != 32:
            raise NotImplementedError("storage='bf16' needs C = 32")
        ops = [m._op for ch in (self.decompation.chain, self.decompation.chain2, self.chain) for m in ch._ops]
        bad = sorted({type(o).__name__ for o in ops if not o.bf16_ok})
        if bad:
            raise NotImplementedError("storage='bf16' covers the primitives of the shipped genotype; no bf16 forward "
                                      "for: %s" % ", ".join(bad))
        return True

    def _engine(self):
        return {'auto': _lib.ENGINE_AUTO, 'direct': _lib.ENGINE_DIRECT, 'tcgen05': _lib.ENGINE_TCGEN05}[self.conv_engine]

    # -------------------------------------------------------------------------------------
    def _run_forward(self, ir, vis, save, capture=None, vis_rgb=False):
        """ir, vis: [B,1,H,W] fp32 CUDA views (any strides); with ``vis_rgb`` vis is the [B,3,H,W] RGB image and the
        stem forms Y on the fly.  Returns (out[B,1,H,W], saved).
        ``capture``: optional dict that receives the intermediate maps ``forward2`` returns (fp32 storage only)."""
        B, _, H, W = ir.shape
        p = self._packed(save)
        bf16 = self._bf16_storage(save)
        rt = _Runtime(B, H, W, self._C, ir.device, _lib.ENGINE_TCGEN05 if bf16 else self._engine(), save, bf16=bf16)
        rt.profile = self.profile
        rt.dilconv_dense = self.dilconv_dense
        C = self._C
        feats, guides, gstats, feats16, gf_ma = [], [], [], [], []
        for img, w, a in ((ir, p["stem_w"][0], p["stem_a"][0]), (vis, p["stem_w"][1], p["stem_a"][1])):
            f, g = rt.new_map(fp32=True), rt.new_plane()
            rt.note_bytes(4 + 4 * C + 4 + (2 * C if bf16 else 0))
            if vis_rgb and img is vis:
                # RGB -> Y inside the stem (RGB2YCrCb, core/model_fusion_auto.py:69-92): no Y plane, no strided view
                f16 = rt.new_map() if bf16 else None
                rt.call("paif_stem_forward_rgb", img.data_ptr(), img.stride(0), img.stride(1), img.stride(2), img.stride(3),
                        w.data_ptr(), a.data_ptr(), f.data_ptr(), g.data_ptr(), _ptr(f16), B, H, W)
                if bf16:
                    feats16.append(f16)
            elif bf16:
                # stems, guide and guided filter stay fp32; the bf16 copy of the stem features is the branch residual
                f16 = rt.new_map()
                rt.call("paif_stem_forward_bf16copy", img.data_ptr(), img.stride(0), img.stride(2), img.stride(3),
                        w.data_ptr(), a.data_ptr(), f.data_ptr(), g.data_ptr(), f16.data_ptr(), B, H, W)
                feats16.append(f16)
            else:
                rt.call("paif_stem_forward", img.data_ptr(), img.stride(0), img.stride(2), img.stride(3),
                        w.data_ptr(), a.data_ptr(), f.data_ptr(), g.data_ptr(), B, H, W)
            feats.append(f)
            guides.append(g)
        d = self.decompation
        branch_out, branch_recs = [], []
        fused_gf = (self.gf_fused and capture is None and rt.tc_engine()
                    and bool(_lib.load().paif_gf_mix_supported(C, H, W)))
        for i, (chain, packs) in enumerate(((d.chain, p["chain_ir"]), (d.chain2, p["chain_vis"]))):
            stats = torch.empty((3, B, H, W), device=ir.device, dtype=torch.float32)
            rt.note_bytes(4 + 12)
            rt.call("paif_gf_guide_stats", guides[i].data_ptr(), stats.data_ptr(), B, H, W)
            if fused_gf:
                x = rt.new_map()
                rt.note_bytes(4 * (C + 4) + (2 if bf16 else 4) * C)
                if save:
                    # also keep mean2(A'): the adjoint's direct guide term, instead of a forward recompute in the backward
                    ma = rt.new_map(fp32=True)
                    rt.note_bytes(4 * (C + 4) + (2 if bf16 else 4) * C + 4 * C)
                    rt.call("paif_gf_mix_forward_save", feats[i].data_ptr(), guides[i].data_ptr(), stats.data_ptr(),
                            p["gfmix_w"][i].data_ptr(), p["c1x1_b"][i].data_ptr(), x.data_ptr(), int(bf16), ma.data_ptr(),
                            C, B, H, W)
                    gf_ma.append(ma)
                else:
                    rt.call("paif_gf_mix_forward", feats[i].data_ptr(), guides[i].data_ptr(), stats.data_ptr(),
                            p["gfmix_w"][i].data_ptr(), p["c1x1_b"][i].data_ptr(), x.data_ptr(), int(bf16), C, B, H, W)
            else:
                lf1, lf2 = rt.new_map(fp32=True), rt.new_map(fp32=True)
                rt.note_bytes(4 * (C + 4) + 8 * C)
                rt.call("paif_gf_decomp_forward", feats[i].data_ptr(), guides[i].data_ptr(), stats.data_ptr(),
                        lf1.data_ptr(), lf2.data_ptr(), C, B, H, W)
                x = rt.conv([lf1, lf2, feats[i]], p["c1x1"][i], ch_shift=p["c1x1_b"][i], src_fp32=True)[0]
                if capture is not None:
                    capture.setdefault("lf", []).append((lf1, lf2))
                del lf1, lf2
                gf_ma.append(None)
            gstats.append(stats if save else None)
            del stats
            o, recs = chain.fwd(rt, packs, x, [feats16[i] if bf16 else feats[i]])
            branch_out.append(o)
            branch_recs.append(recs)
        a_f, v_f = branch_out
        agg = rt.new_map()
        scale = rt.new_plane() if save else None
        rt.note_bytes((2 if bf16 else 4) * 3 * C + (4 if save else 0))
        if bf16:
            rt.call("paif_spa_fused_forward_bf16", p["spa_w"].data_ptr(), p["spa_k"], a_f.data_ptr(), v_f.data_ptr(),
                    agg.data_ptr(), C, B, H, W)
        else:
            rt.call("paif_spa_fused_forward", p["spa_w"].data_ptr(), p["spa_k"], a_f.data_ptr(), v_f.data_ptr(),
                    agg.data_ptr(), _ptr(scale), C, B, H, W)
        f2, recs3 = self.chain.fwd(rt, p["chain"], agg, [])
        out = torch.empty((B, 1, H, W), device=ir.device, dtype=torch.float32)
        pre_out = rt.new_plane() if save else None
        rt.note_bytes((2 if bf16 else 4) * C + 4 + (4 if save else 0))
        if rt.tc_engine() and self.out_tensor_core:
            # interior pixels as an implicit GEMM on the engine, the one-pixel border exactly from the 9-class weights
            rt.call("paif_out_forward_tc", f2.data_ptr(), (p["out_mma16"] if bf16 else p["out_mma"]).data_ptr(),
                    p["out_wm"].data_ptr(), p["out_a"].data_ptr(), out.data_ptr(), _ptr(pre_out),
                    _lib.STORAGE_BF16 if bf16 else _lib.STORAGE_F32, C, B, H, W)
        elif bf16:
            rt.call("paif_out_forward_bf16", f2.data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(),
                    out.data_ptr(), C, B, H, W)
        else:
            rt.call("paif_out_forward", f2.data_ptr(), p["out_wm"].data_ptr(), p["out_a"].data_ptr(), out.data_ptr(),
                    _ptr(pre_out), C, B, H, W)
        self.last_launches = rt.launches
        if capture is not None:
            capture.update(feats=feats, guides=guides, branch_out=branch_out)
        saved = None
        if save:
            saved = dict(B=B, H=H, W=W, feats=feats, guides=guides, gstats=gstats, gf_ma=gf_ma, branch_recs=branch_recs, a_f=a_f, v_f=v_f,
                         scale=scale, recs3=recs3, out=out, pre_out=pre_out, packed=p, bf16=bf16)
        return out, saved

    def _run_backward(self, saved, g):
        """g: [B,1,H,W] contiguous fp32.  Returns (g_ir, g_vis) as [B,H,W] planes."""
        B, H, W, C = saved["B"], saved["H"], saved["W"], self._C
        p = saved["packed"]
        rt = _Runtime(B, H, W, C, g.device, _lib.ENGINE_TCGEN05 if saved.get("bf16") else self._engine(), False)
        rt.profile = self.profile
        rt.dilconv_dense = self.dilconv_dense
        if saved.get("bf16"):
            # The forward ran with bf16 maps (storage='bf16') and saved them as such (half the memory of the fp32 mode
            # until now); the gradient chain itself is fp32 (TF32 operands, fp32 accumulate): widen the saved
            # activations once, and rebuild the attention plane the bf16 blend kernel does not keep.
            saved = dict(saved)
            for key in ("branch_recs", "recs3", "a_f", "v_f"):
                saved[key] = self._widen(rt, saved[key])
            scale, agg_again = rt.new_plane(), rt.new_map()
            rt.call("paif_spa_fused_forward", p["spa_w"].data_ptr(), p["spa_k"], saved["a_f"].data_ptr(),
                    saved["v_f"].data_ptr(), agg_again.data_ptr(), scale.data_ptr(), C, B, H, W)
            saved["scale"] = scale
            del agg_again
        # stem_out + tanh; if the last op of the final chain is a ResidualModule its PReLU' mask is fused here
        gf2 = rt.new_map()
        last = self.chain._ops[-1]._op
        gmasked = None
        if isinstance(last, ResidualModule):
            pre = saved["recs3"][-1][1][0]
            gmasked = rt.new_map()
            rt.call("paif_out_backward", g.data_ptr(), saved["out"].data_ptr(), saved["pre_out"].data_ptr(),
                    p["out_wm"].data_ptr(), p["out_a"].data_ptr(), gf2.data_ptr(), pre.data_ptr(),
                    p["chain"][-1]["a"].data_ptr(), gmasked.data_ptr(), C, B, H, W)
        else:
            rt.call("paif_out_backward", g.data_ptr(), saved["out"].data_ptr(), saved["pre_out"].data_ptr(),
                    p["out_wm"].data_ptr(), p["out_a"].data_ptr(), gf2.data_ptr(), None, None, None, C, B, H, W)
        g_agg = self._chain3_bwd(rt, p, saved, gf2, gmasked)
        a_f, v_f, scale = saved["a_f"], saved["v_f"], saved["scale"]
        gpre = rt.new_plane()
        rt.call("paif_spa_blend_backward_pre", g_agg.data_ptr(), a_f.data_ptr(), v_f.data_ptr(), scale.data_ptr(),
                gpre.data_ptr(), C, B, H, W)
        g_a, g_v = rt.new_map(), rt.new_map()
        rt.call("paif_spa_blend_backward", g_agg.data_ptr(), a_f.data_ptr(), v_f.data_ptr(), scale.data_ptr(),
                gpre.data_ptr(), p["spa_w"].data_ptr(), p["spa_k"], g_a.data_ptr(), g_v.data_ptr(), C, B, H, W)
        d = self.decompation
        grads = []
        for i, (chain, packs, gb) in enumerate(((d.chain, p["chain_ir"], g_a), (d.chain2, p["chain_vis"], g_v))):
            gx = chain.bwd(rt, packs, saved["branch_recs"][i], gb)      # grad w.r.t. the 1x1 conv output
            wd = p["c1x1_d"][i]
            glf1 = rt.conv([gx], wd[0])[0]
            glf2 = rt.conv([gx], wd[1])[0]
            gz = rt.conv([gx], wd[2])[0]
            gfeat = rt.new_map()
            nparts = _lib.load().paif_gf_guide_parts(C)
            gres = torch.empty((nparts, B, H, W), device=g.device, dtype=torch.float32)
            work = torch.empty((_lib.load().paif_gf_backward_work_floats(C, B, H, W),), device=g.device, dtype=torch.float32)
            ma = saved["gf_ma"][i]
            if ma is not None:
                rt.call("paif_gf_decomp_backward_saved", saved["feats"][i].data_ptr(), saved["guides"][i].data_ptr(),
                        saved["gstats"][i].data_ptr(), glf1.data_ptr(), glf2.data_ptr(), gx.data_ptr(), ma.data_ptr(),
                        gfeat.data_ptr(), gres.data_ptr(), work.data_ptr(), C, B, H, W)
            else:
                rt.call("paif_gf_decomp_backward", saved["feats"][i].data_ptr(), saved["guides"][i].data_ptr(),
                        saved["gstats"][i].data_ptr(), glf1.data_ptr(), glf2.data_ptr(), gfeat.data_ptr(), gres.data_ptr(),
                        work.data_ptr(), C, B, H, W)
            del work
            gstem = rt.new_map()
            rt.call("paif_stem_backward_pre", saved["feats"][i].data_ptr(), p["stem_a"][i].data_ptr(),
                    gb.data_ptr(), gz.data_ptr(), gfeat.data_ptr(), None, gres.data_ptr(), nparts, gstem.data_ptr(),
                    C, B, H, W)
            gimg = rt.new_plane()
            rt.call("paif_stem_backward", gstem.data_ptr(), p["stem_w"][i].data_ptr(), gimg.data_ptr(), C, B, H, W)
            grads.append(gimg)
        self.last_launches = rt.launches
        return grads[0], grads[1]

    def _widen(self, rt, obj):
        """bf16 C8 maps inside a saved-activation structure -> fp32 C4 maps (paif_widen_bf16_map); everything else as is."""
        if isinstance(obj, torch.Tensor):
            if obj.dtype == torch.bfloat16 and obj.dim() == 5 and obj.shape[-1] == 8:
                b, o, h, w, _ = obj.shape
                out = torch.empty((b, o * 2, h, w, 4), device=obj.device, dtype=torch.float32)
                rt.call("paif_widen_bf16_map", obj.data_ptr(), out.data_ptr(), o * 8, b, h, w)
                return out
            return obj
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._widen(rt, o) for o in obj)
        return obj

    def _chain3_bwd(self, rt, p, saved, gf2, gmasked):
        chain, packs, recs = self.chain, p["chain"], saved["recs3"]
        if gmasked is None:
            return chain.bwd(rt, packs, recs, gf2)
        # same as Cell_Chain.bwd, with the fused mask handed to the trailing ResidualModule
        gs = gf2
        n = len(chain._ops)
        for i in reversed(range(n)):
            op = chain._ops[i]._op
            inp, rec = recs[i]
            want = [gf2] if i == 0 else []
            fused, rest = want[:op.max_extra_add], want[op.max_extra_add:]
            if i == n - 1:
                gs = op.bwd(rt, packs[i], rec, gs, fused, g_masked=gmasked)
            elif isinstance(op, (DilConv, SepConv)):
                gs = op.bwd(rt, packs[i], rec, gs, fused, x=inp)
            else:
                gs = op.bwd(rt, packs[i], rec, gs, fused)
            if rest:
                gs = rt.add_all(gs, rest)
        return gs

    # -------------------------------------------------------------------------------------
    def _check_inputs(self, ir, vis):
        if self.training:
            raise RuntimeError("paif_b200.Network_Fusion_Searched supports eval() mode only (BatchNorm batch "
                               "statistics are out of scope; the reference scripts always call .eval())")
        if not (ir.is_cuda and vis.is_cuda):
            raise RuntimeError("paif_b200 has no CPU path: inputs must be CUDA tensors")
        if next(self.parameters()).device != ir.device:
            raise RuntimeError("module parameters and inputs live on different devices")
        if ir.dim() != 4 or vis.dim() != 4 or ir.shape[0] != vis.shape[0] or ir.shape[2:] != vis.shape[2:]:
            raise ValueError("expected ir [B,>=1,H,W] and vis [B,>=1,H,W] of the same batch and size")
        if ir.shape[2] <= 9 or ir.shape[3] <= 9:
            raise AssertionError("guided filter (radius 4) needs H, W > 9")

    def forward(self, ir, vis):
        self._check_inputs(ir, vis)
        if ir.dtype != torch.float32:
            ir = ir.float()
        if vis.dtype != torch.float32:
            vis = vis.float()
        with torch.cuda.device(ir.device):
            return _FusionFn.apply(ir, vis, self, False)

    def forward_rgb(self, ir, vis_rgb):
        """``forward(ir, RGB2YCrCb(vis_rgb)[:, 0:1])`` with the RGB -> Y conversion of the task wrappers
        (core/model_fusion_auto.py:69-92, 713-714) done inside the visible stem: ``vis_rgb`` is the [B,3,H,W] RGB image;
        its gradient is the Y-path gradient spread over R, G, B."""
        self._check_inputs(ir, vis_rgb)
        if vis_rgb.shape[1] != 3:
            raise ValueError("forward_rgb expects a [B,3,H,W] RGB visible image")
        with torch.cuda.device(ir.device):
            return _FusionFn.apply(ir.float(), vis_rgb.float(), self, True)

    def _loss(self, ir, vis, mask):
        logits = self(ir, vis)
        return self._criterion(ir, vis, logits, mask)

    def forward2(self, ir, vis):
        """``Network_Fusion_Searched_showfeatures.forward2`` (core/model_fusion_auto.py:669-679): the fused image and
        the intermediate maps the reference exposes for visualisation, as NCHW fp32 tensors:
        ``(output, ir_feature, vis_feature, lf_ir, hf_ir, res_ir, lf_vis, hf_vis, res_vis)`` with
        ``lf_* = cat[LF(eps=1e-3), LF(eps=1e-4)]`` (64 ch), ``hf_* = cat[x - LF1, x - LF2]``, ``res_*`` the
        channel max-min guide (Cell_Decom_decom, :554-585).  Inference only (runs under ``no_grad``, fp32 storage)."""
        if self.storage != 'fp32':
            raise RuntimeError("forward2 returns fp32 intermediates: use storage='fp32'")
        self._check_inputs(ir, vis)
        cap = {}
        with torch.no_grad(), torch.cuda.device(ir.device):
            out, _ = self._run_forward(ir.float()[:, 0:1], vis.float()[:, 0:1], False, capture=cap)

            def nchw(m):                              # C4 map [B][C/4][H][W][4] -> [B][C][H][W]
                b, q, h, w, _ = m.shape
                return m.permute(0, 1, 4, 2, 3).reshape(b, q * 4, h, w)

            res = []
            for i in range(2):
                x = nchw(cap["feats"][i])
                lf = torch.cat([nchw(cap["lf"][i][0]), nchw(cap["lf"][i][1])], 1)
                res.append((lf, torch.cat([x, x], 1) - lf, cap["guides"][i].unsqueeze(1)))
            (lf_ir, hf_ir, res_ir), (lf_vis, hf_vis, res_vis) = res
            return (out, nchw(cap["branch_out"][0]), nchw(cap["branch_out"][1]),
                    lf_ir, hf_ir, res_ir, lf_vis, hf_vis, res_vis)


class _FusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ir, vis, net, vis_rgb=False):
        need = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])   # False under no_grad
        out = None
        if not need and not vis_rgb:
            out, saved = net._run_forward_native(ir[:, 0:1], vis[:, 0:1]), None
        elif need and not vis_rgb:
            res = net._run_forward_native_save(ir[:, 0:1], vis[:, 0:1])
            if res is not None:
                out, saved = res
        if out is None:
            out, saved = net._run_forward(ir[:, 0:1], vis if vis_rgb else vis[:, 0:1], need, vis_rgb=vis_rgb)
        ctx.vis_rgb = vis_rgb
        if saved is not None:
            # `out` is an output of this node: keeping it in a plain attribute would form the cycle
            # out -> grad_fn(ctx) -> saved['out'] -> out and pin ~2 GB of activations per 480x640 frame until the
            # cyclic GC runs (robust_test.py:166 builds such a graph and never calls backward)
            saved.pop("out")
            ctx.save_for_backward(out)
        ctx.net, ctx.saved = net, saved
        ctx.had_saved = saved is not None
        ctx.shapes = (ir.shape, vis.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        net, saved = ctx.net, ctx.saved
        if saved is None:
            if ctx.had_saved:
                raise RuntimeError("Trying to backward through the paif_b200 fusion graph a second time: the saved "
                                   "activations are freed after the first backward (retain_graph is not supported)")
            return None, None, None, None
        g = g.contiguous().float()
        (out,) = ctx.saved_tensors
        with torch.cuda.device(g.device):
            if "native_ws" in saved:
                g_ir, g_vis = net._run_backward_native(saved, g)      # (tanh's output lives inside the workspace)
            else:
                g_ir, g_vis = net._run_backward(dict(saved, out=out), g)
        ctx.saved = None
        outs = []
        for k, (gi, shape, need) in enumerate(((g_ir, ctx.shapes[0], ctx.needs_input_grad[0]),
                                               (g_vis, ctx.shapes[1], ctx.needs_input_grad[1]))):
            if not need:
                outs.append(None)
            elif k == 1 and ctx.vis_rgb:
                # Y = .299 R + .587 G + .114 B: the Y gradient spread over the three colour planes
                # (no host tensor here: this runs inside CUDA-graph capture of the PGD iteration)
                outs.append(torch.stack((gi * 0.299, gi * 0.587, gi * 0.114), dim=1))
            elif shape[1] == 1:
                outs.append(gi.view(shape))
            else:
                full = torch.zeros(shape, device=gi.device, dtype=gi.dtype)
                full[:, 0] = gi
                outs.append(full)
        return outs[0], outs[1], None, None


class Network_Fusion_Searched_showfeatures(Network_Fusion_Searched):
    """core/model_fusion_auto.py:641-697: same parameters and ``forward`` as ``Network_Fusion_Searched`` (its
    ``Cell_Decom_decom`` holds the same sub-modules as ``Cell_Decom``); ``forward2`` additionally returns the
    decomposition intermediates."""

