"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.pt`` by running the
UNMODIFIED reference (``/root/reference``, imported with ``oracle/shims``) on CPU.

Run in the build container:  ``python -m oracle.gen_golden``

Each fixture holds the reference module's ``state_dict`` (45 keys), the inputs, the
reference output ``Network_Fusion_Searched.forward(ir, vis)``
(core/model_fusion_auto.py:625-635), and the reference autograd input-gradients for a
fixed cotangent.  ``vis`` is stored as the contiguous tensor; tests rebuild the
NHWC-strided view the wrappers pass themselves.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import fusion_oracle as fo  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# a second genotype that reaches the other primitives of OPS (operations_m.py:9-18) and other (kernel, dilation)
# pairs through the same constructor: SepConv, SPAattention, 5x5 / dilation-1 / dilation-2 variants
ALT_GENOTYPE = ref_loader.Genotype(
    normal_1=[('SepConv_3_1', 0), ('DilConv_3_1', 1)], normal_1_concat=[1, 2],
    normal_2=[('SPAattention_3', 0), ('Denseblocks_5_1', 1)], normal_2_concat=[1, 2],
    normal_3=[('ECAattention_5', 0), ('Residualblocks_3_2', 1)], normal_3_concat=[1, 2])
GENOTYPES = {"fusion_at": ref_loader.fusion_at, "alt": ALT_GENOTYPE}

CASES = [
    # name, weight seed, randomize, B, H, W, smooth-inputs, genotype
    ("seed0_default_2x40x56", 0, False, 2, 40, 56, False, "fusion_at"),
    ("seed1_random_1x48x72", 1, True, 1, 48, 72, False, "fusion_at"),
    ("seed1_random_smooth_2x33x47", 1, True, 2, 33, 47, True, "fusion_at"),
    ("alt_seed2_random_2x36x52", 2, True, 2, 36, 52, False, "alt"),
]


def make_inputs(B, H, W, smooth, seed=1):
    g = torch.Generator().manual_seed(seed)
    ir = torch.rand(B, 1, H, W, generator=g)
    vis = torch.rand(B, 3, H, W, generator=g)
    if smooth:
        ir = F.avg_pool2d(ir, 9, 1, 4)
        vis = F.avg_pool2d(vis, 9, 1, 4)
    return ir, vis


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = sys.argv[1:]
    for name, seed, rnd, B, H, W, smooth, gname in CASES:
        if only and name not in only:
            continue
        genotype = GENOTYPES[gname]
        net = ref_loader.build_reference_fusion(genotype=genotype, seed=seed, randomize=rnd)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        ir, vis = make_inputs(B, H, W, smooth)
        ir_r = ir.clone().requires_grad_(True)
        vis_r = vis.clone().requires_grad_(True)
        out = net(ir_r, vis_r)
        g = torch.Generator().manual_seed(7)
        gout = torch.randn(out.shape, generator=g)
        g_ir, g_vis = torch.autograd.grad(out, [ir_r, vis_r], gout)
        # validate the restatement against the real reference before writing
        o2, g_ir2, g_vis2 = fo.fusion_input_grads(sd, genotype, ir, vis, gout)
        err = (o2 - out.detach()).abs().max().item()
        gerr = max((g_ir2 - g_ir).abs().max().item(), (g_vis2 - g_vis).abs().max().item())
        print("%-32s out range [%.4f, %.4f]  oracle-vs-reference: out %.3e  grads %.3e"
              % (name, out.min().item(), out.max().item(), err, gerr))
        assert err < 1e-6 and gerr < 1e-4 * max(1.0, g_ir.abs().max().item()), "restatement diverges"
        torch.save({"state_dict": sd, "ir": ir, "vis": vis, "out": out.detach(),
                    "grad_out": gout, "grad_ir": g_ir, "grad_vis": g_vis,
                    "genotype": {f: getattr(genotype, f) for f in genotype._fields},
                    "meta": {"seed": seed, "randomize": rnd, "smooth": smooth, "genotype": gname,
                             "torch": torch.__version__}},
                   os.path.join(GOLDEN, name + ".pt"))


if __name__ == "__main__":
    main()
