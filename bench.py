#!/usr/bin/env python
"""bench.py — PAIF fusion hot path on B200: fused 480x640 pairs/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--batch 16] [--height 480] [--width 640] [--engine auto|direct|tcgen05]

One "step" = one forward pass of Network_Fusion_Searched over one synthetic batch (16 pairs of
480x640 by default: BASELINE.json configs[1] without the stock-PyTorch SegFormer consumer).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

  value   : whole-job pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e     : same metric through the public nn.Module call with HOST (pinned) inputs: H2D of ir and
            vis and D2H of the fused image inside the timed region.
  roofline: the dominant kernel (the dense-convolution engine), algorithmic FLOPs / CUDA-event time
            measured live around every conv launch of the timed steps.
  cpu_baseline / --impl reference: the CPU oracle port (oracle/fusion_oracle.py — the reference's
            operator structure in torch CPU ops) on the box's host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CPU_KIND_NOTE = {"reference": "the UNMODIFIED reference Network_Fusion_Searched from the staged tree baseline/_ref, torch CPU ops",
                 "port": "oracle port, torch CPU ops: no reference tree on this box"}
FLOP_PER_PX = 519336.0          # conv MACs x 2 per pixel of one pair
BYTES_PER_PX = 1891 * 4.0       # layer-granular algorithmic traffic, fp32 storage (1202 ch read + 689 written)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--engine", default="auto", choices=["auto", "direct", "tcgen05"])
    ap.add_argument("--cpu-baseline-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--table", action="store_true", help="print the per-launch CUDA-event table of one step to stderr")
    ap.add_argument("--bwd-steps", type=int, default=3, help="timed forward+backward-to-input steps (0 = skip)")
    ap.add_argument("--bf16-steps", type=int, default=5,
                    help="timed forward steps in the bf16 storage mode (SURVEY 8 config D's precision tier; 0 = skip)")
    ap.add_argument("--eager-steps", type=int, default=3,
                    help="timed steps of the stock-PyTorch eager baseline on the same GPU (rank 0, N=1 only; 0 = skip)")
    ap.add_argument("--pgd-consumer-dtype", default="bf16", choices=["bf16", "tf32"],
                    help="the stock consumer runs under bf16 autocast (default) or in fp32 with TF32 matmul/conv")
    ap.add_argument("--pgd-eager", action="store_true", help="PGD leg without CUDA-graph replay of the PGD iteration")
    ap.add_argument("--pgd-frames", type=int, default=32,
                    help="GLOBAL frames of the PGD-10 robust-eval leg, split over the GPUs (BASELINE configs[2]: "
                         "global batch 32, strong scaling; 0 = skip)")
    ap.add_argument("--pgd-weak-frames", type=int, default=8,
                    help="frames PER GPU of the weak-scaling PGD-10 leg (0 = skip)")
    ap.add_argument("--pgd-micro-batch", type=int, default=8, help="frames attacked together (per-sample min-max wrapper)")
    ap.add_argument("--config-e-batch", type=int, default=32, help="batch of the forward+backward leg (BASELINE configs[4])")
    ap.add_argument("--config-d-steps", type=int, default=3,
                    help="timed steps of the 64x768x1024 fp32 / bf16 sweep (BASELINE configs[3]; N=1 only; 0 = skip)")
    ap.add_argument("--config-b-steps", type=int, default=3,
                    help="timed steps of fusion + SegFormer inference at batch 16 (BASELINE configs[1]; N=1 only; 0 = skip)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def synth_state(seed=0):
    import torch
    import paif_b200
    torch.manual_seed(seed)
    net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
    return net, {k: v.clone() for k, v in net.state_dict().items()}


def synth_inputs(B, H, W, seed=1):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 1, H, W, generator=g), torch.rand(B, 3, H, W, generator=g)


def reference_fusion_module(sd, device="cpu"):
    """The UNMODIFIED reference ``Network_Fusion_Searched`` (core/model_fusion_auto.py:599-640) imported from the
    staged tree (``baseline/_ref`` on the GPU box, /root/reference in the build container) through
    ``oracle/ref_loader`` (import shims for its absent third-party modules; the un-vendored ``guided_filter_pytorch``
    is the restatement in oracle/shims), loaded with the bench's seed-0 weights.  None when no tree is present."""
    import contextlib
    import io
    from oracle import ref_loader
    if not ref_loader.reference_available():
        return None
    m = ref_loader.load_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.Network_Fusion_Searched(32, None, ref_loader.fusion_at)
    net.load_state_dict(sd, strict=True)
    return net.to(device).eval()


def cpu_arm(sd):
    """(kind, forward(ir, vis), input_grads(ir, vis, gout)) of the CPU arm: the reference module itself when its
    tree travelled with the repo (kind "reference"), else the oracle port (kind "port")."""
    import torch
    import paif_b200
    from oracle import fusion_oracle as fo
    ref = reference_fusion_module(sd)
    if ref is None:
        return ("port", lambda ir, vis: fo.fusion_forward(sd, paif_b200.fusion_at, ir, vis),
                lambda ir, vis, g: fo.fusion_input_grads(sd, paif_b200.fusion_at, ir, vis, g))
    for p in ref.parameters():
        p.requires_grad_(False)                       # backward-to-input only, like the drop-in

    def grads(ir, vis, g):
        a, v = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
        return torch.autograd.grad(ref(a, v), [a, v], g)

    return "reference", ref, grads


def cpu_port_fwd_bwd_pairs_per_s(steps, H, W):
    """Forward + backward-to-input of the CPU arm (autograd), one pair per step."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    _, sd = synth_state()
    _, _, grads = cpu_arm(sd)
    ir, vis = synth_inputs(1, H, W)
    gout = torch.rand(1, 1, H, W) - 0.5
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        grads(ir, vis, gout)
        if i >= 1:
            times.append(time.perf_counter() - t0)
    return steps / sum(times)


def cpu_port_pairs_per_s(steps, H, W, warmup=1):
    """The CPU arm on all host cores, one pair per step (bounded sample).  Returns (pairs/s, cores, times, kind)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, sd = synth_state()
    kind, fwd, _ = cpu_arm(sd)
    ir, vis = synth_inputs(1, H, W)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd(ir, vis)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return steps / sum(times), cores, times, kind


def torch_eager_gpu_pairs_per_s(dev, B, H, W, steps=3):
    """SURVEY.md 8d's second reported baseline: the reference's operator structure in STOCK PyTorch CUDA ops (the oracle
    port: F.conv2d / cat / cumsum box filters, fp32, cuDNN with PyTorch's default TF32 setting) on the same B200.
    A baseline leg only — the oracle never runs on the product path."""
    import torch
    import paif_b200
    from oracle import fusion_oracle as fo
    _, sd = synth_state()
    sd = {k: v.to(dev) for k, v in sd.items()}
    ir, vis = synth_inputs(B, H, W)
    ir, vis = ir.to(dev), vis.to(dev)
    with torch.no_grad():
        fo.fusion_forward(sd, paif_b200.fusion_at, ir, vis)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fo.fusion_forward(sd, paif_b200.fusion_at, ir, vis)
        e1.record()
        torch.cuda.synchronize(dev)
    return B * steps / (e0.elapsed_time(e1) * 1e-3)


class SyntheticFrames:
    """Indexable of (vis[3,H,W], ir[1,H,W], label[H,W]) generated from the GLOBAL frame index, so every rank
    sees the same frame for the same index whatever the sharding."""

    def __init__(self, n, H, W, seed=2):
        self.n, self.H, self.W, self.seed = n, H, W, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        import torch
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        return (torch.rand(3, self.H, self.W, generator=g), torch.rand(1, self.H, self.W, generator=g),
                torch.randint(0, 9, (self.H, self.W), generator=g))


def build_consumer(dev, seed=3):
    """The segmentation consumer of the PGD / config-B legs: the reference's OWN SegFormer
    (``core.model_fusion_auto.WeTr('mit_b3', 9, 256)`` = core/mix_transformer.py + core/segformer_head.py, unmodified,
    imported from the staged tree with the timm / mmcv import shims) when the tree travelled with the repo, else the
    from-scratch MiT-B3-shaped stand-in.  Stock PyTorch either way (north_star), random init, parameters frozen."""
    import contextlib
    import io
    import torch
    from oracle import ref_loader
    torch.manual_seed(seed)
    if ref_loader.reference_available():
        m = ref_loader.load_reference()
        with contextlib.redirect_stdout(io.StringIO()):
            seg = m.WeTr('mit_b3', 9, 256, None)
        name = "reference WeTr('mit_b3', 9, 256) (core/mix_transformer.py + core/segformer_head.py, unmodified, staged tree)"
    else:
        from paif_b200.consumer import SegFormerLite
        seg = SegFormerLite(9, 256)
        name = "stand-in SegFormerLite (MiT-B3 shape): no reference tree on this box"
    seg = seg.to(dev).eval()
    for p in seg.parameters():
        p.requires_grad_(False)
    return seg, name


def run_pgd_leg(net, seg, seg_name, args, world, rank, dev, barrier, n_total, tag):
    """BASELINE's second metric: PGD-10 (eps 8/255, alpha 2/255, robust_test.py:40-41) robust-eval frames/s.
    Every frame: 10 x (fusion + consumer forward, backward to the inputs) + 1 attacked forward + confusion update;
    ``n_total`` frames are sharded over the ranks and attacked in micro-batches through the per-sample min-max
    wrapper (= the reference at batch 1 applied to each frame); ONE int64 all-reduce of the 9x9
    confusion matrix at the end (inside the timed region)."""
    import torch
    import torch.distributed as dist
    from paif_b200.consumer import FusionSegTask
    from paif_b200.evaluate import robust_eval, shard_range
    torch.backends.cuda.matmul.allow_tf32 = True            # the stock consumer may use TF32 matmuls
    torch.backends.cudnn.allow_tf32 = True
    cast = None if args.pgd_consumer_dtype == "tf32" else torch.bfloat16
    task = FusionSegTask(net, seg, consumer_autocast=cast, per_sample_minmax=True, fused_glue=True).to(dev).eval()
    H, W = args.height, args.width
    per_gpu = len(shard_range(n_total, 0, world))
    mb = max(1, min(args.pgd_micro_batch, per_gpu))
    graphed = not args.pgd_eager
    warm = SyntheticFrames(n_total + mb * world, H, W)
    robust_eval(task, [warm[n_total + rank * mb + i] for i in range(mb)], attack_iters=2, use_cuda_graph=graphed,
                micro_batch=mb)                                  # warm-up: allocator, autotune, graph capture
    barrier()
    # fusion-only share: forward + backward-to-input of the fusion net alone at the micro-batch size
    gout = torch.rand(mb, 1, H, W, device=dev) - 0.5
    xi, xv = torch.rand(mb, 1, H, W, device=dev), torch.rand(mb, 1, H, W, device=dev)

    def fusion_fb():
        a, v = xi.detach().requires_grad_(True), xv.detach().requires_grad_(True)
        net(a, v).backward(gout)

    for _ in range(2):
        fusion_fb()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(3):
        fusion_fb()
    f1.record()
    torch.cuda.synchronize()
    fusion_fb_ms = f0.elapsed_time(f1) / 3
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    meter = robust_eval(task, SyntheticFrames(n_total, H, W), attack_iters=10, rank=rank, world_size=world,
                        use_cuda_graph=graphed, micro_batch=mb)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    conf = meter.conf.cpu()
    batches = (per_gpu + mb - 1) // mb
    fusion_ms = batches * 10.5 * fusion_fb_ms                  # 10 fwd+bwd + one forward (~ half a fwd+bwd) per micro-batch
    return {"metric": "PGD-10 robust-eval frames/s", "value": n_total / (ms * 1e-3), "unit": "frames/s",
            "scaling": tag, "frames": n_total, "frames_per_gpu": per_gpu, "micro_batch": mb,
            "ms_per_frame_per_gpu": ms / max(per_gpu, 1),
            "fusion_share": {"fusion_fwd_bwd_ms_per_micro_batch": fusion_fb_ms, "estimated_fusion_ms": fusion_ms,
                             "fraction_of_leg": fusion_ms / ms,
                             "note": "fusion net forward+backward-to-input timed alone at the micro-batch size x "
                                     "(10 + 0.5) per micro-batch; the rest is the stock consumer, the loss head and the delta update"},
            "attack": "PGD-10 eps 8/255 alpha 2/255, l_seg loss, seeded start per global frame index",
            "consumer": "%s, %s, params frozen" % (seg_name, "fp32 with TF32 matmul/conv" if cast is None else "bf16 autocast"),
            "cuda_graph": graphed,
            "confusion_sum": int(conf.sum()), "confusion_all_reduce": "int64 SUM over %d rank(s)" % world,
            "confusion_trace": int(conf.diag().sum())}


def measure_tf32_peak(dev, seconds=1.5):
    """cuBLAS TF32 8192^3 the way MEASURED_PEAKS.json measures bf16: best of 10 (burst) and back to back (sustained)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        return {"burst_tflops": flops / (best * 1e-3) / 1e12, "sustained_tflops": flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": "torch.matmul fp32 inputs with allow_tf32 (cuBLAS TF32) 8192^3: best of 10, and %d back to back" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    v, cores, times, kind = cpu_port_pairs_per_s(args.steps, args.height, args.width, warmup=min(args.warmup, 1))
    sample = "%d timed forward passes of 1 pair %dx%d (fp32, no_grad, %s) after %d warm-up" % (
        args.steps, args.height, args.width, CPU_KIND_NOTE[kind], min(args.warmup, 1))
    line = {"impl": "reference", "metric": "fused %dx%d pairs/s (fusion-net forward)" % (args.height, args.width),
            "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Network_Fusion_Searched forward, fusion_at genotype, random-init (seed 0), batch %d "
                                   "synthetic %dx%d IR+RGB pairs per GPU (BASELINE configs[1] without the stock-PyTorch "
                                   "SegFormer consumer)" % (args.batch, args.height, args.width),
                       "batch_per_gpu": args.batch, "height": args.height, "width": args.width,
                       "note": "CPU arm: each step is a bounded sample of that workload (1 pair of the batch)"},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import paif_b200
    from paif_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner (and any NCCL_DEBUG output) to
        # fd 1 while the communicator is created, so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    B, H, W = args.batch, args.height, args.width
    net, _ = synth_state()
    net = net.to(dev).eval()
    net.conv_engine = args.engine
    ir_h, vis_h = synth_inputs(B, H, W, seed=1 + rank)
    ir_h, vis_h = ir_h.pin_memory(), vis_h.pin_memory()
    out_h = torch.empty(B, 1, H, W).pin_memory()
    ir_d, vis_d = ir_h.to(dev), vis_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def timed_e2e(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e2e_drain()                                          # the last D2H is part of the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def step_resident():
        with torch.no_grad():
            return net(ir_d, vis_d)

    # end to end: host buffers in, host buffer out, through the public nn.Module call.  The H2D copies of step
    # i+1 and the D2H copy of step i run on their own streams (the two PCIe directions are independent) while the
    # kernels of step i run on the compute stream (two device input slots); every step's copies are issued inside
    # the timed region.
    copy_stream = torch.cuda.Stream(device=dev)          # host -> device
    d2h_stream = torch.cuda.Stream(device=dev)           # device -> host

    def probe_gbs(dst, src, stream):
        """one timed pinned copy: reported next to e2e so that a slow PCIe path on a box is visible in the line"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            dst.copy_(src, non_blocking=True)
            e0.record(stream)
            dst.copy_(src, non_blocking=True)
            e1.record(stream)
        e1.synchronize()
        return src.numel() * src.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9

    h2d_gbs = probe_gbs(torch.empty_like(vis_d), vis_h, copy_stream)
    d2h_gbs = probe_gbs(out_h, torch.empty(B, 1, H, W, device=dev), d2h_stream)
    slots = [(torch.empty_like(ir_d), torch.empty_like(vis_d)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]      # inputs of slot landed
    freed = [torch.cuda.Event() for _ in range(2)]      # kernels that read slot finished
    state = {"i": 0, "primed": False}

    def issue_h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            slots[slot][0].copy_(ir_h, non_blocking=True)
            slots[slot][1].copy_(vis_h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["i"] & 1
        if not state["primed"]:
            freed[0].record(); freed[1].record()
            issue_h2d(cur)
            state["primed"] = True
        issue_h2d(cur ^ 1)                                   # next step's inputs, overlapped with this step's kernels
        main = torch.cuda.current_stream(dev)
        main.wait_event(ready[cur])
        with torch.no_grad():
            out = net(slots[cur][0], slots[cur][1])
        freed[cur].record(main)
        done = torch.cuda.Event()
        done.record(main)
        out.record_stream(d2h_stream)                        # keep the allocator from recycling it under the copy
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            out_h.copy_(out, non_blocking=True)              # D2H of the fused images
        state["i"] += 1

    def e2e_drain():
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        torch.cuda.current_stream(dev).wait_stream(d2h_stream)

    # per-step working set (> 30 fp32 maps of B*39 MB) is far larger than the 126 MB L2
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step_resident, args.steps)                    # the headline: no per-launch events, the module as a user calls it
    launches = net.last_launches * args.steps
    net.profile = []                                         # the same steps again with CUDA events around every launch:
    ms_prof = timed(step_resident, args.steps)               # the per-kernel roofline table (per-operator path of the module)
    prof, net.profile = net.profile, None
    def e2e_steps():
        step_e2e()
    ms_e2e = timed_e2e(e2e_steps, args.steps)

    # attack inner step (BASELINE configs[4]): forward + backward-to-input on resident inputs
    fwd_bwd = None
    if args.bwd_steps > 0:
        Be = args.config_e_batch
        ir_e, vis_e = synth_inputs(Be, H, W, seed=11 + rank)
        ir_e, vis_e = ir_e.to(dev), vis_e.to(dev)
        gout = torch.rand(Be, 1, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) - 0.5

        def step_fwd_bwd():
            a = ir_e.detach().requires_grad_(True)
            v = vis_e.detach().requires_grad_(True)
            net(a, v).backward(gout)
            return a.grad, v.grad

        for _ in range(2):
            step_fwd_bwd()
        net.profile = [] if args.table else None
        ms_fb = timed(step_fwd_bwd, args.bwd_steps)
        prof_fb, net.profile = net.profile, None
        fb_gbs = 2.0 * BYTES_PER_PX * H * W * Be * args.bwd_steps / (ms_fb * 1e-3) / 1e9
        fwd_bwd = {"value": Be * world * args.bwd_steps / (ms_fb * 1e-3), "unit": "pairs/s", "ms_per_step": ms_fb / args.bwd_steps,
                   "steps": args.bwd_steps, "launches_per_step": net.last_launches, "batch_per_gpu": Be,
                   "roofline": {"bound": "hbm", "achieved": fb_gbs, "peak": peaks()["hbm_gbs"], "unit": "GB/s",
                                "frac": fb_gbs / peaks()["hbm_gbs"],
                                "note": "algorithmic bytes = 2 x the forward's layer-granular plan (SURVEY 8d: 1891 ch x 4 B "
                                        "per pixel; every backward layer reads its output gradient and writes its input "
                                        "gradient once, mirroring the forward) / step time / measured HBM peak"},
                   "what": "BASELINE configs[4]: forward + backward-to-input (the PGD inner step of attack/attack.py:444-501 "
                           "without the segmentation consumer), batch %d per GPU at %dx%d, inputs resident" % (Be, H, W)}
        del ir_e, vis_e, gout
        if args.table and rank == 0 and prof_fb:
            per = len(prof_fb) // args.bwd_steps
            agg = {}
            for (n, m, a, b) in prof_fb[-per:]:
                key = n if not (m and "k" in m) else "%s k%d d%d cin%d" % (n, m["k"], m["dil"], m["cin"])
                t = a.elapsed_time(b)
                agg[key] = (agg.get(key, (0, 0.0))[0] + 1, agg.get(key, (0, 0.0))[1] + t)
            tot = sum(t for _, t in agg.values())
            sys.stderr.write("--- forward+backward step, launches aggregated by kind ---\n")
            for key, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                sys.stderr.write("%-44s x%-3d %8.3f ms %5.1f%%\n" % (key, cnt, t, 100 * t / tot))
            sys.stderr.write("sum of launches %.3f ms (fwd+bwd step %.3f ms)\n" % (tot, ms_fb / args.bwd_steps))

    # bf16 storage mode (north_star's 1e-2 tier): the same forward step with bf16 activation maps / bf16 tensor-core
    # operands after the decomposition; reported beside the headline, never instead of it
    bf16_leg = None
    if args.bf16_steps > 0 and args.engine != "direct":
        with torch.no_grad():
            out32 = net(ir_d, vis_d)
        net.storage = 'bf16'
        try:
            for _ in range(3):
                step_resident()
            net.profile = [] if args.table else None
            ms16 = timed(step_resident, args.bf16_steps)
            prof16, net.profile = net.profile, None
            with torch.no_grad():
                err16 = (net(ir_d, vis_d) - out32).abs().max().item()
            bf16_leg = {"value": B * world * args.bf16_steps / (ms16 * 1e-3), "unit": "pairs/s",
                        "ms_per_step": ms16 / args.bf16_steps, "steps": args.bf16_steps,
                        "launches_per_step": net.last_launches, "dtype": "bf16 maps + bf16 MMA operands, fp32 accumulate; "
                        "stems / guided filter / decomposition inputs fp32",
                        "max_abs_vs_fp32_mode": err16, "tolerance": 1e-2,
                        "whole_step_frac_of_hbm_roof": (BYTES_PER_PX / 2 * H * W * B * args.bf16_steps / (ms16 * 1e-3) / 1e9) / peaks()["hbm_gbs"],
                        "whole_step_note": "SURVEY 8d bf16 plan (1891 ch x 2 B per pixel) / step time / HBM peak",
                        "what": "the headline forward step with net.storage='bf16', batch %d per GPU, inputs resident" % B}
            if args.table and rank == 0 and prof16:
                per = len(prof16) // args.bf16_steps
                tot = sum(a.elapsed_time(b) for (_, _, a, b) in prof16[-per:])
                sys.stderr.write("--- forward step, bf16 storage mode ---\n")
                for (n, m, a, b) in prof16[-per:]:
                    t = a.elapsed_time(b)
                    extra = ""
                    if m and "k" in m:
                        extra = " k%d d%d cin%d  %.1f TFLOP/s  %.0f GB/s" % (m["k"], m["dil"], m["cin"], m["flops"] / (t * 1e-3) / 1e12,
                                                                             m["bytes"] / (t * 1e-3) / 1e9)
                    sys.stderr.write("%-28s %8.3f ms %5.1f%%%s\n" % (n, t, 100 * t / tot, extra))
                sys.stderr.write("sum of launches %.3f ms (step %.3f ms)\n" % (tot, ms16 / args.bf16_steps))
        finally:
            net.storage = 'fp32'
            net.profile = None

    if args.table and rank == 0:
        per = len(prof) // args.steps
        tot = sum(a.elapsed_time(b) for (_, _, a, b) in prof[-per:])
        for (n, m, a, b) in prof[-per:]:
            t = a.elapsed_time(b)
            extra = ""
            if m and "k" in m:
                extra = " k%d d%d cin%d  %.1f TFLOP/s  %.0f GB/s" % (m["k"], m["dil"], m["cin"], m["flops"] / (t * 1e-3) / 1e12,
                                                                     m["bytes"] / (t * 1e-3) / 1e9)
            sys.stderr.write("%-28s %8.3f ms %5.1f%%%s\n" % (n, t, 100 * t / tot, extra))
        sys.stderr.write("sum of launches %.3f ms (step %.3f ms)\n" % (tot, ms / args.steps))
    torch.cuda.empty_cache()
    pgd = pgd_weak = None
    seg = seg_name = None
    if args.pgd_frames > 0 or args.pgd_weak_frames > 0 or (args.config_b_steps > 0 and world == 1):
        seg, seg_name = build_consumer(dev)
    if args.pgd_frames > 0:
        pgd = run_pgd_leg(net, seg, seg_name, args, world, rank, dev, barrier, max(args.pgd_frames, world), "strong")
    if args.pgd_weak_frames > 0:
        pgd_weak = run_pgd_leg(net, seg, seg_name, args, world, rank, dev, barrier, args.pgd_weak_frames * world, "weak")
    torch.cuda.empty_cache()

    # BASELINE configs[1] as written: fusion + SegFormer inference at batch 16 (N = 1 only; the consumer is stock PyTorch)
    config_b = None
    if args.config_b_steps > 0 and world == 1 and seg is not None:
        from paif_b200.consumer import FusionSegTask
        cast = None if args.pgd_consumer_dtype == "tf32" else torch.bfloat16
        task = FusionSegTask(net, seg, consumer_autocast=cast, per_sample_minmax=True, fused_glue=True).to(dev).eval()

        def step_b():
            with torch.no_grad():
                return task(ir_d, vis_d)

        for _ in range(2):
            step_b()
        ms_b = timed(step_b, args.config_b_steps)
        config_b = {"value": B * args.config_b_steps / (ms_b * 1e-3), "unit": "pairs/s", "ms_per_step": ms_b / args.config_b_steps,
                    "fusion_ms_per_step": ms / args.steps, "batch": B,
                    "what": "BASELINE configs[1] (test_original.py path): colour glue + fusion drop-in + SegFormer inference, "
                            "batch %d synthetic %dx%d, 1 GPU" % (B, H, W),
                    "consumer": "%s, %s" % (seg_name, "fp32 with TF32 matmul/conv" if cast is None else "bf16 autocast")}
        del task
    seg = None
    torch.cuda.empty_cache()

    # BASELINE configs[3]: fusion-only sweep at the M3FD shape 768x1024, batch 64, fp32 storage vs bf16 storage, with a
    # max-abs check of one sample of each against the CPU arm (N = 1 only)
    config_d = None
    if args.config_d_steps > 0 and world == 1 and args.engine != "direct":
        Bd, Hd, Wd = 64, 768, 1024
        ir_b, vis_b = synth_inputs(Bd, Hd, Wd, seed=21)
        ir_b, vis_b = ir_b.to(dev), vis_b.to(dev)
        config_d = {"what": "BASELINE configs[3]: fusion-only forward at %dx%d, batch %d, fp32 (TF32 MMA) vs bf16 storage" % (Hd, Wd, Bd)}
        outs = {}
        for mode in ("fp32", "bf16"):
            net.storage = mode

            def step_d():
                with torch.no_grad():
                    return net(ir_b, vis_b)

            for _ in range(2):
                step_d()
            ms_d = timed(step_d, args.config_d_steps)
            outs[mode] = step_d()[0:1].cpu()
            bpp = BYTES_PER_PX if mode == "fp32" else BYTES_PER_PX / 2
            config_d[mode] = {"value": Bd * args.config_d_steps / (ms_d * 1e-3), "unit": "pairs/s",
                              "ms_per_step": ms_d / args.config_d_steps,
                              "whole_step_frac_of_hbm_roof": (bpp * Hd * Wd * Bd * args.config_d_steps / (ms_d * 1e-3) / 1e9) / peaks()["hbm_gbs"]}
        net.storage = 'fp32'
        if not args.no_cpu_baseline:
            _, sd_ = synth_state()
            kind_, fwd_, _g = cpu_arm(sd_)
            torch.set_num_threads(os.cpu_count() or 1)
            with torch.no_grad():
                ref_d = fwd_(ir_b[0:1].cpu(), vis_b[0:1].cpu())
            config_d["parity"] = {"cpu_arm": kind_, "sample": "pair 0 of the timed batch",
                                  "max_abs_fp32_storage": (outs["fp32"] - ref_d).abs().max().item(), "tolerance_fp32": 1e-3,
                                  "max_abs_bf16_storage": (outs["bf16"] - ref_d).abs().max().item(), "tolerance_bf16": 1e-2}
        del ir_b, vis_b, outs
        torch.cuda.empty_cache()
    tf32_peak = measure_tf32_peak(dev) if (rank == 0 and world == 1) else None
    clocks = sampler.summary() if sampler else None          # sampled across every timed GPU leg above

    # dominant kernel: the dense-conv engine (all conv launches of the timed steps).  After the wide-N MMA
    # rewrite every conv shape of the genotype except the 7x7 is HBM-bound, so the engine is judged against
    # the HBM roofline: algorithmic bytes (each source / residual map read once, each output written once).
    conv = [(m, a.elapsed_time(b)) for (n, m, a, b) in prof if n == "paif_conv_forward"]
    other_ms = sum(a.elapsed_time(b) for (n, m, a, b) in prof if n != "paif_conv_forward")
    conv_ms = sum(t for _, t in conv)
    conv_flops = sum(m["flops"] for m, _ in conv)
    conv_bytes = sum(m["bytes"] for m, _ in conv)
    pk = peaks()
    engine_id = conv[0][0]["engine"] if conv else 0
    peak_tf = tf32_peak["sustained_tflops"] if tf32_peak else pk["bf16_tflops_sustained"] / 2.0
    # every launch of the step against its own roofline (algorithmic bytes / flops from the call site, CUDA-event time)
    per = len(prof) // max(args.steps, 1)
    agg = {}
    for i, (n, m, a, b) in enumerate(prof):
        key = (i % per, n)
        e = agg.setdefault(key, {"name": n, "ms": 0.0, "bytes": (m or {}).get("bytes"), "flops": (m or {}).get("flops"),
                                 "shape": ("k%d d%d cin%d" % (m["k"], m["dil"], m["cin"])) if (m and "k" in m) else None})
        e["ms"] += a.elapsed_time(b) / max(args.steps, 1)
    per_kernel = []
    for key in sorted(agg):
        e = agg[key]
        if e["bytes"]:
            gbs = e["bytes"] / (e["ms"] * 1e-3) / 1e9
            e["gbs"], e["frac_hbm"] = gbs, gbs / pk["hbm_gbs"]
        if e["flops"]:
            e["tflops"] = e["flops"] / (e["ms"] * 1e-3) / 1e12
            e["frac_tf32"] = e["tflops"] / peak_tf
        per_kernel.append({k: v for k, v in e.items() if v is not None})
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    achieved_gbs = conv_bytes / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")          # from the committed ncu --set full capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pairs = B * world * args.steps
    value = pairs / (ms * 1e-3)
    line = {
        "metric": "fused %dx%d pairs/s (fusion-net forward)" % (H, W), "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if engine_id != _lib.ENGINE_TCGEN05 else "tf32",
        "data": "synthetic",
        "config": {"workload": "Network_Fusion_Searched forward, fusion_at genotype, random-init (seed 0), batch %d "
                               "synthetic %dx%d IR+RGB pairs per GPU (BASELINE configs[1] without the stock-PyTorch "
                               "SegFormer consumer)" % (B, H, W),
                   "batch_per_gpu": B, "height": H, "width": W, "conv_engine": {1: "direct-fp32", 2: "tcgen05-tf32"}.get(engine_id, "?"),
                   "l2": "per-step working set (>30 fp32 maps x %.0f MB) exceeds the 126 MB L2; no flush needed" % (B * H * W * 128 / 1e6),
                   "parallelism": "dp%d (independent batches, no data-path collective)" % world},
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(ir_h.numel() * 4 + vis_h.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4),
                "pinned": bool(ir_h.is_pinned() and vis_h.is_pinned() and out_h.is_pinned()),
                "h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs,
                "note": "H2D of step i+1 and D2H of step i overlap the kernels of step i (separate copy streams); "
                        "h2d_gbs / d2h_gbs = this box's pinned PCIe copy rates"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "dense-conv engine conv_tc_kernel (all %d conv launches per step)" % (len(conv) // max(args.steps, 1)),
                     "achieved": achieved_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / pk["hbm_gbs"],
                     "traffic": traffic, "traffic_source": "static: profiles/conv_traffic.json (ncu --set full capture of the "
                                                           "conv launches; not re-measured by this run)",
                     "peak_note": "%s HBM copy bandwidth (MEASURED_PEAKS.json)" % pk["source"],
                     "bytes_per_launch": conv_bytes / max(len(conv), 1), "avg_launch_ms": conv_ms / max(len(conv), 1),
                     "profiled_ms_per_step": ms_prof / args.steps,
                     "profiled_note": "launch durations come from a second pass over the same steps with CUDA events around "
                                      "every launch (the per-operator path of the module); `value` is timed without them",
                     "conv_share_of_step": conv_ms / (conv_ms + other_ms) if conv_ms + other_ms > 0 else None,
                     "tensor_tflops": achieved_tf, "tensor_frac_of_tf32_peak": achieved_tf / peak_tf,
                     "tf32_peak_note": ("cuBLAS TF32 8192^3 measured by this run (sustained)" if tf32_peak else
                                        "TF32 = 1/2 x %s sustained bf16 cuBLAS peak" % pk["source"]),
                     "per_kernel": per_kernel,
                     "whole_step_frac_of_hbm_roof": (BYTES_PER_PX * H * W * B * args.steps / (ms * 1e-3) / 1e9) / pk["hbm_gbs"],
                     "whole_step_note": "SURVEY 8d layer-granular bytes (1891 ch x 4 B per pixel) / step time / HBM peak"},
        "clocks": clocks,
    }
    if fwd_bwd is not None:
        line["fwd_bwd"] = fwd_bwd
    if bf16_leg is not None:
        line["bf16_storage"] = bf16_leg
    if args.eager_steps > 0 and world == 1:
        eb = min(B, 4)                                       # eager keeps ~25 map-sized temporaries per guided filter
        try:
            torch.cuda.empty_cache()
            line["torch_eager_same_gpu"] = {
                "value": torch_eager_gpu_pairs_per_s(dev, eb, H, W, args.eager_steps), "unit": "pairs/s", "batch": eb,
                "what": "the reference's operator structure in stock PyTorch CUDA ops (oracle port, fp32, no_grad) on "
                        "this GPU: the eager implementation a user of the reference runs today; reported, not a target"}
        except Exception as ex:                              # a baseline leg must never take the bench line down
            line["torch_eager_same_gpu"] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    if pgd is not None:
        line["pgd10"] = pgd
    if pgd_weak is not None:
        line["pgd10_weak"] = pgd_weak
    if config_b is not None:
        line["config_b"] = config_b
    if config_d is not None:
        line["config_d"] = config_d
    if tf32_peak is not None:
        line["tf32_peak"] = tf32_peak
    if not args.no_cpu_baseline and world == 1:
        v, cores, times, kind = cpu_port_pairs_per_s(args.cpu_baseline_steps, H, W)
        st = sorted(times)
        line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind,
                                "sample": "%d timed forward passes of 1 pair %dx%d on the host (%s) after 1 warm-up"
                                          % (args.cpu_baseline_steps, H, W, CPU_KIND_NOTE[kind]),
                                "min_ms": 1e3 * st[0], "median_ms": 1e3 * st[len(st) // 2],
                                "fwd_bwd_pairs_per_s": cpu_port_fwd_bwd_pairs_per_s(2, H, W),
                                "fwd_bwd_sample": "2 timed forward+backward-to-input passes of 1 pair (autograd of the same module) after 1 warm-up"}
        _, sd0 = synth_state()
        kind0, fwd0, _g0 = cpu_arm(sd0)
        with torch.no_grad():
            ref0 = fwd0(ir_h[0:1], vis_h[0:1])
            got0 = net(ir_d[0:1].contiguous(), vis_d[0:1].contiguous()).cpu()
            got16 = net(ir_d, vis_d)[0:1].cpu()
        line["parity"] = {"cpu_arm": kind0, "what": "pair 0 of the timed batch: the drop-in's output (alone and inside the timed "
                          "batch of %d) against the CPU arm on the same weights and inputs" % B,
                          "max_abs_batch1": (got0 - ref0).abs().max().item(), "max_abs_in_timed_batch": (got16 - ref0).abs().max().item(),
                          "tolerance": 1e-3 if engine_id == _lib.ENGINE_TCGEN05 else 5e-5}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
