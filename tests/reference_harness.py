"""TEST INFRASTRUCTURE ONLY: runs the reference's own wrappers, attack and entry scripts UNCHANGED around the
paif_b200 drop-in (north_star: "test_original.py and robust_test.py run unchanged").

The reference tree is taken from ``oracle/ref_loader.reference_root()``: /root/reference in the build container,
the staged copy ``baseline/_ref`` (``__graft_entry__.stage_reference``) on the GPU box.  What the scripts hard-code
and this harness supplies without touching them ("unchanged-script plumbing"):

* the seven third-party modules the reference imports but the image lacks -> ``oracle/shims`` on ``sys.path``;
* dataset directories ``/user33/objectdetection/test_all/{Visible,Infrared,Label}/`` (robust_test.py:100-102,
  test_original.py:102-104) -> ``TaskFusion_dataset2.prepare_data_path`` (the directory lister) is wrapped so that
  this prefix maps to a synthetic PNG dataset in the work directory (the script still constructs and iterates
  ``Fusion_dataset('val', ...)`` itself);
* ``./checkpoint/model_meta30000_fusion_8.pth`` (robust_test.py:259), ``./model_Proposed_wodenfense_fusion_best.pth``
  (test_original.py:715), ``pretrained/mit_b3.pth`` (core/model_fusion_auto.py:23), ``configs/voc.yaml``
  (robust_test.py:29) -> random-init files written into the work directory, which becomes the cwd;
* the fusion class -> ``paif_b200.install()`` rebinds ``core.model_fusion_auto.Network_Fusion_Searched`` before
  the script's ``from core.model_fusion_auto import ...`` runs (``use_dropin=True``), or leaves the reference's own
  class in place (``use_dropin=False``: the run the drop-in run is compared with).
"""
import contextlib
import io
import os
import runpy
import shutil
import sys

import numpy as np
import torch

from paif_testutil import ROOT  # noqa: F401  (puts the repo root on sys.path)
from oracle import ref_loader

DATA_PREFIX = "/user33/objectdetection/test_all/"


def available():
    return ref_loader.reference_available()


def reference_modules():
    """(core.model_fusion_auto, attack.attack) of the unmodified reference, shims installed."""
    m = ref_loader.load_reference()
    import attack.attack as atk
    return m, atk


def make_dataset(root, n_frames, H, W, seed=0):
    """Synthetic MFNet-shaped PNG triples (RGB visible, 8-bit infrared, 9-class label) under ``root``."""
    from PIL import Image
    rng = np.random.RandomState(seed)
    for sub in ("Visible", "Infrared", "Label"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n_frames):
        name = "%05d.png" % i
        base = (127 + 90 * np.sin(xx / (7.0 + i) + i) * np.cos(yy / (5.0 + 2 * i))).astype(np.float32)
        vis = np.clip(base[..., None] + rng.randint(-40, 40, (H, W, 3)), 0, 255).astype(np.uint8)
        ir = np.clip(255 - base + rng.randint(-30, 30, (H, W)), 0, 255).astype(np.uint8)
        label = ((xx // max(W // 6, 1) + yy // max(H // 4, 1) + i) % 9).astype(np.uint8)
        Image.fromarray(vis).save(os.path.join(root, "Visible", name))
        Image.fromarray(ir).save(os.path.join(root, "Infrared", name))
        Image.fromarray(label).save(os.path.join(root, "Label", name))


def build_reference_task(kind, seed=0, backbone="mit_b3"):
    """The reference's task model around its OWN fusion net, random init under ``seed``:
    kind 'searched' -> ``Network_MM_Searched`` (core/model_fusion_auto.py:1029-1060, robust_test.py:262),
    kind 'comp'     -> ``Network_MM_CompModel`` (:698-729, test_original.py:716-719)."""
    m, _ = reference_modules()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "searched":
            net = m.Network_MM_Searched(32, ref_loader.fusion_at, None, None, backbone, num_classes=9)
        else:
            fusion = m.Network_Fusion_Searched(32, None, ref_loader.fusion_at)
            net = m.Network_MM_CompModel(fusion, None, None, backbone, 9, 256, None)
    ref_loader.randomize_state(net.enhance_net, seed)
    return net.eval()


def build_dropin_task(kind, state_dict, backbone="mit_b3"):
    """The same reference wrapper class around the paif_b200 fusion net, loaded ``strict=True`` from the
    reference-built model's ``state_dict``."""
    import paif_b200
    m, _ = reference_modules()
    saved = m.Network_Fusion_Searched
    try:
        paif_b200.install(m)
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "searched":
                net = m.Network_MM_Searched(32, ref_loader.fusion_at, None, None, backbone, num_classes=9)
            else:
                fusion = m.Network_Fusion_Searched(32, None, ref_loader.fusion_at)
                net = m.Network_MM_CompModel(fusion, None, None, backbone, 9, 256, None)
    finally:
        m.Network_Fusion_Searched = saved
    assert isinstance(net.enhance_net, paif_b200.Network_Fusion_Searched)
    net.load_state_dict(state_dict, strict=True)
    return net.eval()


def prepare_workdir(workdir, script, n_frames=2, H=96, W=128, seed=0):
    """Everything ``script`` ('robust_test.py' | 'test_original.py') expects to find relative to its cwd."""
    ref = ref_loader.reference_root()
    os.makedirs(workdir, exist_ok=True)
    make_dataset(os.path.join(workdir, "test_all"), n_frames, H, W, seed)
    os.makedirs(os.path.join(workdir, "configs"), exist_ok=True)
    shutil.copy(os.path.join(ref, "configs", "voc.yaml"), os.path.join(workdir, "configs", "voc.yaml"))
    if script == "robust_test.py":
        net = build_reference_task("searched", seed)
        os.makedirs(os.path.join(workdir, "checkpoint"), exist_ok=True)
        torch.save(net.state_dict(), os.path.join(workdir, "checkpoint", "model_meta30000_fusion_8.pth"))
    elif script == "test_original.py":
        net = build_reference_task("comp", seed)
        torch.save(net.state_dict(), os.path.join(workdir, "model_Proposed_wodenfense_fusion_best.pth"))
        # WeTr(..., pretrained=True) loads pretrained/mit_b3.pth and pops the ImageNet head (core/model_fusion_auto.py:22-26)
        enc = {k: v.clone() for k, v in net.denoise_net.encoder.state_dict().items()}
        enc["head.weight"], enc["head.bias"] = torch.zeros(1000, 512), torch.zeros(1000)
        os.makedirs(os.path.join(workdir, "pretrained"), exist_ok=True)
        torch.save(enc, os.path.join(workdir, "pretrained", "mit_b3.pth"))
    else:
        raise ValueError(script)
    return workdir


def run_script(script, workdir, argv=(), use_dropin=True, seed=0):
    """``runpy`` the UNMODIFIED reference script with cwd = ``workdir``.  Returns its captured stdout."""
    import paif_b200
    ref = ref_loader.reference_root()
    m, _ = reference_modules()
    import TaskFusion_dataset2 as ds
    orig_lister, orig_fusion = ds.prepare_data_path, m.Network_Fusion_Searched
    orig_show = getattr(m, "Network_Fusion_Searched_showfeatures", None)

    def redirected_lister(path):
        if path is not None and path.startswith(DATA_PREFIX):
            path = os.path.join(workdir, "test_all", path[len(DATA_PREFIX):])
        return orig_lister(path)

    cwd, old_argv = os.getcwd(), sys.argv
    out = io.StringIO()
    try:
        ds.prepare_data_path = redirected_lister
        if use_dropin:
            paif_b200.install(m)
        os.chdir(workdir)
        sys.argv = [script] + list(argv)
        torch.manual_seed(seed)                      # the scripts' PGD start point comes from the global RNG
        with contextlib.redirect_stdout(out):
            runpy.run_path(os.path.join(ref, script), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = old_argv
        ds.prepare_data_path = orig_lister
        m.Network_Fusion_Searched = orig_fusion
        if orig_show is not None:
            m.Network_Fusion_Searched_showfeatures = orig_show
    return out.getvalue()


def read_pngs(directory):
    from PIL import Image
    return {n: np.asarray(Image.open(os.path.join(directory, n))) for n in sorted(os.listdir(directory))}
