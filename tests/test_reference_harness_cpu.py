"""CPU check of the unchanged-script harness (tests/reference_harness.py): the reference's test_original.py runs
via ``runpy`` in a synthetic work directory with the reference's own fusion class (the drop-in has no CPU path; the
drop-in runs of the same harness are the ``-m gpu`` tests in test_gpu_reference_dropin.py).  The scripts hard-code
``.cuda()``; on this GPU-less container those calls are neutralised for the duration of the test."""
import os

import pytest
import torch

import reference_harness as rh

pytestmark = pytest.mark.skipif(not rh.available() or torch.cuda.is_available(),
                                reason="needs the reference tree and a GPU-less host")


def test_test_original_runs_unchanged_in_the_harness(tmp_path, monkeypatch):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    wd = rh.prepare_workdir(str(tmp_path / "wd"), "test_original.py", n_frames=1, H=64, W=96)
    out = rh.run_script("test_original.py", wd, ["--num_workers", "0", "--gpu", "-1"], use_dropin=False)
    assert "model load done" in out
    root = os.path.join(wd, "attack", "our_orignal_l_seg_PGD5_8_2_both")
    assert os.path.isfile(os.path.join(root, "our_orignal_PGD5_8_2.txt"))
    fused = rh.read_pngs(os.path.join(root, "fused_attacked"))
    assert list(fused) == ["00000.png"] and fused["00000.png"].shape == (64, 96, 3)
    import core.model_fusion_auto as m
    import TaskFusion_dataset2 as ds
    assert m.Network_Fusion_Searched.__module__ == "core.model_fusion_auto"      # patches are undone
    assert ds.prepare_data_path.__module__ == "TaskFusion_dataset2"
