"""CPU: host logic of the sharded evaluation (paif_b200/evaluate.py) — partitioning, seeded PGD start,
metric conventions, and the world_size-2 integer all-reduce over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fusion_oracle as fo
from paif_b200 import evaluate as ev


def test_shard_range_is_a_contiguous_partition():
    for n in (0, 1, 7, 32, 33, 1000):
        for world in (1, 2, 3, 4, 8):
            parts = [list(ev.shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_seeded_delta_depends_only_on_seed_and_global_index():
    a = ev.seeded_delta((1, 3, 8, 9), 8 / 255., seed=5, global_index=17, device="cpu")
    b = ev.seeded_delta((1, 3, 8, 9), 8 / 255., seed=5, global_index=17, device="cpu")
    c = ev.seeded_delta((1, 3, 8, 9), 8 / 255., seed=5, global_index=18, device="cpu")
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert a.abs().max().item() <= 8 / 255.


def _reference_style_results(conf):
    """util/util.py:31-55 restated with the reference's loop structure (consider_unlabeled=True)."""
    n = conf.shape[0]
    p, r, i = np.zeros(n), np.zeros(n), np.zeros(n)
    for cid in range(n):
        p[cid] = np.nan if conf[:, cid].sum() == 0 else float(conf[cid, cid]) / float(conf[:, cid].sum())
        r[cid] = np.nan if conf[cid, :].sum() == 0 else float(conf[cid, cid]) / float(conf[cid, :].sum())
        den = conf[cid, :].sum() + conf[:, cid].sum() - conf[cid, cid]
        i[cid] = np.nan if den == 0 else float(conf[cid, cid]) / float(den)
    return p, r, i


def test_compute_results_follows_reference_conventions():
    g = torch.Generator().manual_seed(3)
    conf = torch.randint(0, 1000, (9, 9), generator=g, dtype=torch.int64)
    conf[:, 4] = 0          # a class never predicted  -> precision NaN
    conf[6, :] = 0          # a class never present    -> recall NaN
    got = ev.compute_results(conf)
    want = _reference_style_results(conf.numpy())
    for a, b in zip(got, want):
        np.testing.assert_allclose(a.numpy(), b, rtol=0, atol=0, equal_nan=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11)
    labels = torch.randint(0, 10, (n_frames, 24, 31), generator=g)      # 9 = out of range, ignored
    labels[labels == 9] = 255
    preds = torch.randint(0, 9, (n_frames, 24, 31), generator=g)
    conf = torch.zeros(9, 9, dtype=torch.int64)
    for i in ev.shard_range(n_frames, rank, world):
        conf += fo.confusion_matrix(labels[i], preds[i], 9)
    ev.all_reduce_confusion(conf)
    torch.save(conf, os.path.join(out_dir, "conf_%d.pt" % rank))
    dist.destroy_process_group()


def test_world2_gloo_all_reduce_is_bit_exact(tmp_path):
    n_frames, world = 7, 2
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(11)
    labels = torch.randint(0, 10, (n_frames, 24, 31), generator=g)
    labels[labels == 9] = 255
    preds = torch.randint(0, 9, (n_frames, 24, 31), generator=g)
    full = fo.confusion_matrix(labels, preds, 9)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), "conf_%d.pt" % r))
        assert got.dtype == torch.int64 and torch.equal(got, full)
    assert int(full.sum()) == int((labels != 255).sum())


def test_task_glue_matches_reference_colour_conventions_and_consumer_shapes():
    """FusionSegTask restates Network_MM_CompModel.forward's glue (core/model_fusion_auto.py:69-111, 712-729);
    checked here against a direct evaluation of the reference's formulas, with a trivial fusion stand-in."""
    import torch.nn as nn
    from paif_b200.consumer import FusionSegTask, SegFormerLite

    class MeanFusion(nn.Module):
        def forward(self, ir, vis):
            return 0.5 * (ir + vis)

    torch.manual_seed(0)
    task = FusionSegTask(MeanFusion(), SegFormerLite(9, 256, depths=(1, 1, 1, 1))).eval()
    ir, vis = torch.rand(2, 1, 64, 96), torch.rand(2, 3, 64, 96)
    fused, seg = task(ir, vis)
    assert fused.shape == (2, 1, 64, 96) and seg.shape == (2, 9, 16, 24)
    # reference formulas
    R, G, B = vis[:, 0], vis[:, 1], vis[:, 2]
    Y = 0.299 * R + 0.587 * G + 0.114 * B
    Cr, Cb = (R - Y) * 0.713 + 0.5, (B - Y) * 0.564 + 0.5
    torch.testing.assert_close(fused[:, 0], 0.5 * (ir[:, 0] + Y))
    flat = torch.stack([fused[:, 0], Cr, Cb], -1).reshape(-1, 3)
    mat = torch.tensor([[1.0, 1.0, 1.0], [1.403, -0.714, 0.0], [0.0, -0.344, 1.773]])
    rgb = (flat + torch.tensor([0.0, -0.5, -0.5])).mm(mat).reshape(2, 64, 96, 3).permute(0, 3, 1, 2).clamp(0, 1)
    rgb = (rgb - rgb.min()) / (rgb.max() - rgb.min())
    x = (rgb * 255 - task.mean) / task.std
    torch.testing.assert_close(task.denoise_net(x), seg, rtol=1e-4, atol=1e-4)


def test_pgd_delta_order_is_the_references():
    """``attack_both`` returns ``(delta_ir, delta_vis)`` (attack/attack.py:514) and robust_test.py:145 swaps
    the names at the call site; paif_b200 returns a named tuple in that order.  The statement is pinned against the
    reference source when the tree is present (build container / staged baseline/_ref)."""
    assert ev.PGDDelta._fields == ("delta_ir", "delta_vis")
    from oracle import ref_loader
    if ref_loader.reference_available():
        src = open(os.path.join(ref_loader.reference_root(), "attack", "attack.py")).read()
        body = src[src.index("def attack_both("):]
        assert "return delta_ir, delta_vis" in body[:body.index("\ndef ", 1)]


def test_pack_cache_is_invalidated_by_load_state_dict_and_apply():
    """stale packed weights.  ``load_state_dict`` / ``.to()`` / ``invalidate_packed()`` drop the pack and
    change ``pack_signature`` (what a captured PGD graph compares before replaying)."""
    import paif_b200
    torch.manual_seed(0)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at).eval()
    net._pack_cache = ("sentinel", {})
    sig0 = net.pack_signature()
    net.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    assert net._pack_cache is None and net.pack_signature() != sig0
    net._pack_cache = ("sentinel", {})
    sig1 = net.pack_signature()
    net.double()
    assert net._pack_cache is None and net.pack_signature() != sig1
    net.float()
    sig2 = net.pack_signature()
    with torch.no_grad():
        net.stem_1[1].weight.data.mul_(2.0)          # .data edits do not bump _version: explicit invalidation
    assert net.pack_signature() == sig2
    net.invalidate_packed()
    assert net.pack_signature() != sig2
    net.conv_engine = 'direct'
    assert net.pack_signature() != sig2
