"""TEST INFRASTRUCTURE ONLY — CPU restatement of PAIF's fusion hot path.

A functional (state_dict-driven) restatement of
``Network_Fusion_Searched.forward`` (reference core/model_fusion_auto.py:625-635) and
everything it reaches, written with plain torch CPU ops in the *same operator
structure* as the reference (``F.conv2d``, ``torch.cat``, cumsum box filter), so it
doubles as the "port" CPU baseline.  It is differentiable, so
``torch.autograd.grad`` on it is the oracle for the backward-to-input
(attack/attack.py:501).

Pinning: the reference ships no tests or golden vectors.  This
restatement is pinned instead against the reference module itself, imported
unmodified in the build container (``oracle/gen_golden.py`` →
``tests/golden/*.pt``; ``tests/test_oracle.py`` re-checks the fixtures on every
run).  The guided filter comes from an un-vendored, un-pinned third-party package
(``guided_filter_pytorch``): for that dependency parity is UNPINNED — the published
algorithm is restated in ``box_filter`` / ``guided_filter`` below.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
import this file.
"""
import torch
import torch.nn.functional as F

EPS_LIST = (0.001, 0.0001)   # core/model_fusion_auto.py:498
RADIUS = 4                   # core/model_fusion_auto.py:497


# --------------------------------------------------------------------------- #
# guided_filter_pytorch (third party, un-vendored): published algorithm
# --------------------------------------------------------------------------- #
def _diff(c, r, dim):
    n = c.shape[dim]
    left = c.narrow(dim, r, r + 1)
    middle = c.narrow(dim, 2 * r + 1, n - 2 * r - 1) - c.narrow(dim, 0, n - 2 * r - 1)
    right = c.narrow(dim, n - 1, 1) - c.narrow(dim, n - 2 * r - 1, r)
    return torch.cat([left, middle, right], dim=dim)


def box_filter(x, r=RADIUS):
    """(2r+1)^2 window sum clipped to the image (BoxFilter of DeepGuidedFilter)."""
    return _diff(_diff(x.cumsum(dim=2), r, 2).cumsum(dim=3), r, 3)


def guided_filter(x, y, r, eps):
    """GuidedFilter(r, eps)(x, y): x = guide (1 ch), y = source."""
    h, w = x.shape[2:]
    assert h > 2 * r + 1 and w > 2 * r + 1
    N = box_filter(x.new_ones((1, 1, h, w)), r)
    mean_x = box_filter(x, r) / N
    mean_y = box_filter(y, r) / N
    cov_xy = box_filter(x * y, r) / N - mean_x * mean_y
    var_x = box_filter(x * x, r) / N - mean_x * mean_x
    A = cov_xy / (var_x + eps)
    b = mean_y - A * mean_x
    mean_A = box_filter(A, r) / N
    mean_b = box_filter(b, r) / N
    return mean_A * x + mean_b


# --------------------------------------------------------------------------- #
# operations_m.py primitives
# --------------------------------------------------------------------------- #
def basicconv_padding(k, d):
    """BasicConv padding table, operations_m.py:121-132 (anything else → 0)."""
    return {(3, 1): 1, (3, 2): 2, (5, 1): 2, (5, 2): 4, (7, 1): 3, (7, 2): 6}.get((k, d), 0)


def prelu(x, w):
    return F.prelu(x, w)


def residual_dense_block(sd, p, x, k, d):
    """ResidualDenseBlock.forward, operations_m.py:444-449."""
    pad = basicconv_padding(k, d)
    a = sd[p + 'lrelu.weight']
    x1 = prelu(F.conv2d(x, sd[p + 'conv1.conv.weight'], None, 1, pad, d), a)
    x2 = prelu(F.conv2d(torch.cat((x, x1), 1), sd[p + 'conv2.conv.weight'], None, 1, pad, d), a)
    x3 = prelu(F.conv2d(torch.cat((x, x1, x2), 1), sd[p + 'conv3.conv.weight'], None, 1, pad, d), a)
    return x3 * 0.333333 + x


def _bn_eval(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + 'running_mean'], sd[p + 'running_var'],
                        sd[p + 'weight'], sd[p + 'bias'], False, 0.0, eps)


def dil_conv(sd, p, x, k, d):
    """DilConv.forward, operations_m.py:494-506 (eval-mode BN)."""
    pad = basicconv_padding(k, d)
    C = x.shape[1]
    t = F.relu(x)
    t = F.conv2d(t, sd[p + 'op.1.conv.weight'], None, 1, pad, d, groups=C)
    t = F.conv2d(t, sd[p + 'op.2.weight'])
    t = _bn_eval(t, sd, p + 'op.3.')
    return t + x


def eca_basic_block(sd, p, x, k):
    """ECABasicBlock.forward operations_m.py:382-393 + eca_layer.forward :352-366.
    conv1 is conv3x3 (:283), conv2 = BasicConv(k, dilation 1) (:374)."""
    a = sd[p + 'relu.weight']
    x0 = F.conv2d(x, sd[p + 'conv1.weight'], None, 1, 1)
    out = prelu(x0, a)
    out = F.conv2d(out, sd[p + 'conv2.conv.weight'], None, 1, basicconv_padding(k, 1))
    y = out.mean(dim=(2, 3), keepdim=True)                       # AdaptiveAvgPool2d(1)
    y = F.conv1d(y.squeeze(-1).transpose(-1, -2), sd[p + 'se.conv.weight'], None, 1, (k - 1) // 2)
    y = torch.sigmoid(y.transpose(-1, -2).unsqueeze(-1))
    out = out * y.expand_as(out)
    out = out + x0
    return prelu(out, a)


def residual_module(sd, p, x, k, d):
    """ResidualModule.forward, operations_m.py:451-464 (eval-mode BN)."""
    t = F.conv2d(x, sd[p + 'op.0.conv.weight'], None, 1, basicconv_padding(k, d), d)
    t = F.conv2d(t, sd[p + 'op.1.weight'], None, 1, 2, 2)
    t = F.conv2d(t, sd[p + 'op.2.weight'])
    t = _bn_eval(t, sd, p + 'op.3.')
    t = prelu(t, sd[p + 'op.4.weight'])
    return x + t


def sep_conv(sd, p, x, k):
    """SepConv.forward, operations_m.py:509-526 (padding k//2, dilation 1, eval-mode BN; OPS drops the dilation)."""
    C = x.shape[1]
    t = F.relu(x)
    t = F.conv2d(t, sd[p + 'op.1.weight'], None, 1, k // 2, 1, groups=C)
    t = F.conv2d(t, sd[p + 'op.2.weight'])
    t = _bn_eval(t, sd, p + 'op.3.')
    t = F.relu(t)
    t = F.conv2d(t, sd[p + 'op.5.weight'], None, 1, k // 2, 1, groups=C)
    t = F.conv2d(t, sd[p + 'op.6.weight'])
    return _bn_eval(t, sd, p + 'op.7.')


def spatial_basic_block(sd, p, x, k):
    """Spatial_BasicBlock.forward operations_m.py:193-206 (with_norm=False) + the 1-arg ChannelPool /
    spatial_attn_layer of operations_m.py:146-163 (operations_m's own ChannelPool, not the 2-arg override)."""
    a = sd[p + 'relu.weight']
    x0 = F.conv2d(x, sd[p + 'conv1.weight'], None, 1, 1)
    out = prelu(x0, a)
    out = F.conv2d(out, sd[p + 'conv2.conv.weight'], None, 1, basicconv_padding(k, 1))
    pool = torch.cat((out.max(1)[0].unsqueeze(1), out.mean(1).unsqueeze(1)), dim=1)
    scale = torch.sigmoid(F.conv2d(pool, sd[p + 'se.spatial.conv.weight'], None, 1, basicconv_padding(k, 1)))
    return prelu(out * scale + x0, a)


def parse_primitive(primitive):
    """MixedOp name grammar, core/model_fusion_auto.py:404-410."""
    parts = primitive.split('_')
    if primitive.find('attention') != -1:
        return parts[0], int(parts[1]), 1
    return parts[0], int(parts[1]), int(parts[2])


def mixed_op(sd, p, primitive, x):
    name, k, d = parse_primitive(primitive)
    p = p + '_op.'
    if name == 'Denseblocks':
        return residual_dense_block(sd, p, x, k, d)
    if name == 'DilConv':
        return dil_conv(sd, p, x, k, d)
    if name == 'ECAattention':
        return eca_basic_block(sd, p, x, k)
    if name == 'Residualblocks':
        return residual_module(sd, p, x, k, d)
    if name == 'SepConv':
        return sep_conv(sd, p, x, k)
    if name == 'SPAattention':
        return spatial_basic_block(sd, p, x, k)
    raise NotImplementedError(primitive)


def cell_chain(sd, p, types, x):
    """Cell_Chain.forward, core/model_fusion_auto.py:439-445 (sequential; indices unused)."""
    s = x
    for i, (primitive, _idx) in enumerate(types):
        s = mixed_op(sd, '%s_ops.%d.' % (p, i), primitive, s)
    return x + s


# --------------------------------------------------------------------------- #
# Cell_Decom and the network
# --------------------------------------------------------------------------- #
def get_residue(t):
    """core/model_fusion_auto.py:517-521."""
    return torch.max(t, dim=1, keepdim=True)[0] - torch.min(t, dim=1, keepdim=True)[0]


def decomposition(x):
    """core/model_fusion_auto.py:522-535: LF = cat over eps, HF = x - LF."""
    res = get_residue(x)
    LF, HF = [], []
    for eps in EPS_LIST:
        lf = guided_filter(res, x, RADIUS, eps)
        LF.append(lf)
        HF.append(x - lf)
    return torch.cat(LF, 1), torch.cat(HF, 1)


def cell_decom(sd, genotype, fir, fvis, inter=None):
    """Cell_Decom.forward, core/model_fusion_auto.py:509-516."""
    p = 'decompation.'
    lf_ir, hf_ir = decomposition(fir)
    lf_vis, hf_vis = decomposition(fvis)
    lf = F.conv2d(torch.cat([lf_ir, hf_ir], 1), sd[p + 'conv1x1_lf.weight'], sd[p + 'conv1x1_lf.bias'])
    hf = F.conv2d(torch.cat([lf_vis, hf_vis], 1), sd[p + 'conv1x1_hf.weight'], sd[p + 'conv1x1_hf.bias'])
    lf_re = cell_chain(sd, p + 'chain.', genotype.normal_1, lf)
    hf_re = cell_chain(sd, p + 'chain2.', genotype.normal_2, hf)
    if inter is not None:
        inter.update(lf_ir=lf_ir, lf_vis=lf_vis, lf=lf, hf=hf, lf_re=lf_re, hf_re=hf_re)
    return lf_re + fir, hf_re + fvis


def spatial_attn(sd, ir_f, vis_f):
    """2-arg ChannelPool + spatial_attn_layer_M, core/model_fusion_auto.py:1352-1368."""
    pool = torch.cat((ir_f.max(1)[0].unsqueeze(1), ir_f.mean(1).unsqueeze(1),
                      vis_f.max(1)[0].unsqueeze(1), vis_f.mean(1).unsqueeze(1)), dim=1)
    w = sd['spa.spatial.conv.weight']
    k = w.shape[-1]
    return torch.sigmoid(F.conv2d(pool, w, None, 1, basicconv_padding(k, 1)))


def fusion_forward(sd, genotype, ir, vis, inter=None):
    """Network_Fusion_Searched.forward, core/model_fusion_auto.py:625-635.

    ``sd``: the module's ``state_dict`` (45 keys for ``fusion_at``); ``inter``: optional
    dict that receives named intermediates for per-kernel parity tests."""
    vis = vis[:, 0:1]
    ir = ir[:, 0:1]
    fir = prelu(F.conv2d(ir, sd['stem_1.0.weight'], None, 1, 1), sd['stem_1.1.weight'])
    fvis = prelu(F.conv2d(vis, sd['stem_2.0.weight'], None, 1, 1), sd['stem_2.1.weight'])
    ir_f, vis_f = cell_decom(sd, genotype, fir, fvis, inter)
    scale = spatial_attn(sd, ir_f, vis_f)
    agg = scale * ir_f + (1 - scale) * vis_f
    feature2 = cell_chain(sd, 'chain.', genotype.normal_3, agg)
    t = F.conv2d(feature2, sd['stem_out.0.weight'], None, 1, 1)
    t = F.conv2d(t, sd['stem_out.1.weight'], None, 1, 1)
    out = torch.tanh(prelu(t, sd['stem_out.2.weight']))
    if inter is not None:
        inter.update(fir=fir, fvis=fvis, ir_feature=ir_f, vis_feature=vis_f, scale=scale,
                     agg=agg, feature2=feature2, out=out)
    return out


def fusion_input_grads(sd, genotype, ir, vis, grad_out):
    """Oracle for the backward-to-input (autograd of the restatement; the reference
    obtains the same quantity through ``loss.backward()``, attack/attack.py:501)."""
    ir = ir.detach().clone().requires_grad_(True)
    vis = vis.detach().clone().requires_grad_(True)
    out = fusion_forward(sd, genotype, ir, vis)
    g_ir, g_vis = torch.autograd.grad(out, [ir, vis], grad_out)
    return out.detach(), g_ir, g_vis


def confusion_matrix(label, pred, num_classes=9):
    """sklearn.metrics.confusion_matrix(labels=0..n-1) as used at robust_test.py:207-211
    (rows = true label, cols = prediction; labels outside 0..n-1 are ignored)."""
    label = label.reshape(-1).long()
    pred = pred.reshape(-1).long()
    ok = (label >= 0) & (label < num_classes) & (pred >= 0) & (pred < num_classes)
    idx = label[ok] * num_classes + pred[ok]
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)
