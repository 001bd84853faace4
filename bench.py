#!/usr/bin/env python
"""bench.py -- images/s of RRNet's post-backbone path (decode -> stage-1 NMS -> RoIAlign+ReLU ->
head -> generate_bbox) on synthetic 1088x1920 VisDrone-shaped inputs (BASELINE.json configs[1]:
batch 8, 10x272x480 heat-map, K=1500, 256-channel stride-4 features).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N>1; weak scaling: every rank runs the config-2 batch on its own
images, then the detections are all-gathered over NCCL for mAP, inside the timed region).
Prints ONE JSON line on rank 0.  Keys beyond the base contract:
  roofline      dominant kernel: algorithmic bytes per launch / CUDA-event duration of that kernel
                inside the timed region, against MEASURED_PEAKS.json (hbm_gbs)
  cpu_baseline  the reference's CPU call sequence (oracle/ref_port.py: torch CPU + torchvision CPU)
                timed on this box's host cores on a bounded sample of the same workload
  e2e           same metric through the public API with HOST (pinned) buffers: H2D of every
                input and D2H of the result inside the timed region
  stages_ms     mean per-stage device time from events recorded by the library between stages
`--impl reference` times the reference arm only (CPU, all host threads, bounded sample per step).
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

WORKLOAD = dict(B=8, C=10, H=272, W=480, K=1500, feat_ch=256)
WORKLOAD_NAME = "rrnet_eval_post_backbone_1088x1920_b8_k1500"
METRIC = "images/sec post-backbone decode+RoIAlign+NMS at 1088x1920"
UNIT = "images/s"


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference(steps, warmup, sample_images=1, seed=None):
    """Time the reference's CPU call sequence (oracle/ref_port.py) on `sample_images` images of the
    config-2 workload per step, with every host thread torch can use.  Returns (img/s, info)."""
    from oracle import ref_port
    from rrnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = WORKLOAD
    x = synth.eval_inputs(sample_images, w["H"], w["W"], w["K"], synth.SEED_C2 if seed is None else seed)
    hp = synth.head_params(synth.SEED_C2)
    stage = {}
    for _ in range(warmup):
        ref_port.post_backbone(x["hm"], x["wh"], x["off"], x["feat"], hp, w["K"])
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_port.post_backbone(x["hm"], x["wh"], x["off"], x["feat"], hp, w["K"], stage_times=stage)
    dt = time.perf_counter() - t0
    ips = sample_images * steps / dt
    info = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d image(s) of the config-2 workload per step (10x272x480 heat-map, K=%d, 256x272x480 "
                      "features), %d steps after %d warm-up; torch %s CPU + torchvision CPU call sequence of "
                      "models/rrnet.py (oracle/ref_port.py), %d threads" % (
                          sample_images, w["K"], steps, warmup, torch.__version__, cores),
            "ms_per_image": 1e3 * dt / (steps * sample_images),
            "stage_ms_per_image": {k: 1e3 * v / (steps * sample_images) for k, v in stage.items()}}
    return ips, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, max(args.warmup, 1)
    ips, info = cpu_reference(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": info["ms_per_image"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "step": "1 image (bounded sample of the batch-8 step)", **WORKLOAD},
            "cpu_baseline": info,
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cuda_reference(x_dev, hp_dev, K, reps=5):
    """The reference's own GPU path for the same batch: its unmodified call sequence (torch sigmoid/topk/
    gather, torchvision.ops.nms per class with the host syncs that implies, torch.relu + torchvision
    roi_align, the Bottleneck head through cuDNN) on CUDA tensors -- oracle/ref_port.post_backbone.
    This is the baseline BASELINE.json's ">=10x" target names.  Returns ms per batch."""
    from oracle import ref_port
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for _ in range(2):
            ref_port.post_backbone(x_dev["hm"], x_dev["wh"], x_dev["off"], x_dev["feat"], hp_dev, K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ref_port.post_backbone(x_dev["hm"], x_dev["wh"], x_dev["off"], x_dev["feat"], hp_dev, K)
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / reps
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def aux_kernels(dev, peak, feat=None, rois=None):
    """The other rows of the scope table, timed alone (CUDA events, L2 flushed between repetitions):
    training side at config 3 (B=32, 10x128x128) and the config-5 NMS stress (20k boxes, one class)."""
    from rrnet_b200 import ops, synth
    # L2 flush by READING 256 MB (a write flush would leave 126 MB of dirty lines whose write-back is then
    # billed to the kernel under test)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    sink = torch.zeros(1, dtype=torch.float32, device=dev)

    def do_flush():
        torch.sum(flush, dim=0, keepdim=True, out=sink)

    def timed(fn, reps=10):
        """Device time of one fn() with a cold L2: a CUDA graph of reps x (flush L2, fn) minus a graph of
        reps x (flush L2), so that neither Python / ctypes launch overhead nor the flush is counted."""
        fn(); fn(); do_flush()
        torch.cuda.synchronize()

        def graph_ms(with_fn):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    do_flush()
                    if with_fn:
                        fn()
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); b.synchronize()
            return a.elapsed_time(b)

        try:
            return max(graph_ms(True) - graph_ms(False), 0.0) / reps
        except Exception:                                # not capturable: eager launches (includes launch gaps)
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(reps):
                do_flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); b.synchronize()
                tot += a.elapsed_time(b)
            return tot / reps

    out = {}
    B, C, h, w = 32, 10, 128, 128
    g = torch.Generator().manual_seed(synth.SEED_C3)
    z = (torch.randn(B, C, h, w, generator=g) * 2 - 2).to(dev)
    annos, n_obj = synth.pad_annos(synth.train_annos(B, 512, 512, synth.SEED_C3))
    annos, n_obj = annos.to(dev), n_obj.to(dev)
    gt = ops.render_targets(annos, n_obj, 512, 512)[0]
    n = z.numel()
    ms = timed(lambda: ops.focal_fwd_bwd(z, gt))
    out["focal_fwd_bwd_c3"] = {"ms": ms, "algorithmic_bytes": 3 * n * 4, "gbs": 3 * n * 4 / ms / 1e6,
                               "frac_of_hbm_peak": 3 * n * 4 / ms / 1e6 / peak}
    ms = timed(lambda: ops.focal_forward(z, gt))
    out["focal_fwd_c3"] = {"ms": ms, "algorithmic_bytes": 2 * n * 4, "gbs": 2 * n * 4 / ms / 1e6,
                           "frac_of_hbm_peak": 2 * n * 4 / ms / 1e6 / peak}
    ms = timed(lambda: ops.render_targets(annos, n_obj, 512, 512))
    rb = annos.numel() * 4 + n * 4
    out["render_targets_c3"] = {"ms": ms, "algorithmic_bytes": rb, "gbs": rb / ms / 1e6, "frac_of_hbm_peak": rb / ms / 1e6 / peak,
                                "objects": int(n_obj.sum())}
    def fused():
        st = ops.focal_render_forward(z, annos, n_obj, 512, 512)
        return ops.focal_render_backward(z, annos, n_obj, 512, 512, st)
    ms = timed(fused)
    fb = 2 * n * 4 + annos.numel() * 4            # logits read once (second pass from L2) + gradient write
    out["focal_render_fused_fwd_bwd_c3"] = {"ms": ms, "algorithmic_bytes": fb, "gbs": fb / ms / 1e6,
                                            "frac_of_hbm_peak": fb / ms / 1e6 / peak,
                                            "replaces": "render_targets + focal_fwd_bwd (84 MB of traffic)"}
    # RegL1Loss of a wh map at config 3 (B=32, 2x128x128, <=150 objects per image): ours vs the reference's formula
    # (permute copy of the map + gather + l1_loss + autograd) on the same device
    whm = torch.randn(B, 2, h, w, generator=g).to(dev)
    tgt = ops.render_targets(annos, n_obj, 512, 512)
    t_wh, t_ind, t_mask = tgt[1], tgt[2], tgt[4]
    ms = timed(lambda: ops.regl1_fwd_bwd(whm, t_mask, t_ind, t_wh))

    def regl1_reference():
        o = whm.detach().requires_grad_(True)
        pred = o.permute(0, 2, 3, 1).contiguous().view(B, -1, 2).gather(1, t_ind.long().expand(B, t_ind.shape[1], 2))
        m = t_mask.expand_as(pred).float()
        (torch.nn.functional.l1_loss(pred * m, t_wh * m, reduction="sum") / (m.sum() + 1e-4)).backward()
        return o.grad
    ms_ref = timed(regl1_reference, reps=5)
    out["regl1_fwd_bwd_c3"] = {"ms": ms, "reference_formula_torch_cuda_ms": ms_ref,
                               "algorithmic_bytes": int(t_wh.numel() * 4 * 2 + whm.numel() * 4),
                               "note": "gradient map write (4.2 MB memset) dominates; the reference moves the map 4x more"}
    # stage-2 regression loss (the per-image loop of criterion) at a config-2-like size: 8 images x 1400 RoIs x 150 GT rows
    nb, nr, mg = 8, 1400, 150
    gxy = torch.rand(nb, mg, 2, generator=g) * 1500
    gtb = torch.cat([gxy, gxy + torch.rand(nb, mg, 2, generator=g) * 100 + 8, torch.zeros(nb, mg, 4)], 2).to(dev)
    pick = torch.randint(0, mg, (nb, nr), generator=g)
    bx = torch.gather(gtb[..., :4].cpu(), 1, pick[..., None].expand(nb, nr, 4)) + (torch.rand(nb, nr, 4, generator=g) - 0.5) * 10
    bxy = torch.cat([torch.arange(nb).float().repeat_interleave(nr)[:, None], bx.reshape(-1, 4) / 4], 1).to(dev)
    sreg = (torch.randn(nb * nr, 4, generator=g) * 0.5).to(dev)
    sseg = (torch.arange(nb + 1) * nr).int().to(dev)
    ms = timed(lambda: ops.stage2_loss(bxy, sseg, sreg, gtb, 4.0))
    out["stage2_loss_8x1400"] = {"ms": ms, "pairs": nb * nr * mg,
                                 "replaces": "8 x (box_iou, max, mask-select, generate_bbox_target, smooth_l1) with 2 host syncs per image"}
    # RoIAlign backward at the bench workload (config 2 features, the step's own RoIs): gather kernel vs autograd
    # through torchvision.ops.roi_align(relu(feat)) on the same GPU (atomic scatter + ReLU backward)
    if feat is not None and rois is not None and rois.shape[0] > 0:
        gout = torch.randn(rois.shape[0], feat.shape[1], 3, 3, generator=g).to(dev)
        rws = torch.empty(ops._lib.lib().rr_roi_align_workspace_bytes(rois.shape[0], *feat.shape), dtype=torch.uint8, device=dev)
        ms = timed(lambda: ops.roi_align_backward(feat, rois, gout, ws=rws), reps=3)
        entry = {"ms": ms, "rois": int(rois.shape[0]), "algorithmic_bytes": int(2 * feat.numel() * 4 + gout.numel() * 4),
                 "gbs": (2 * feat.numel() * 4 + gout.numel() * 4) / ms / 1e6}
        try:
            import torchvision

            def tv_backward():
                f = feat.detach().requires_grad_(True)
                o = torchvision.ops.roi_align(torch.relu(f), rois, (3, 3))
                return torch.autograd.grad(o, f, gout)[0]
            tv_backward()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                tv_backward()
            b.record(); b.synchronize()
            entry["torchvision_cuda_fwd_bwd_ms"] = a.elapsed_time(b) / 3
            f = feat.detach().requires_grad_(True)
            tv_out = torchvision.ops.roi_align(torch.relu(f), rois, (3, 3))
            torch.cuda.synchronize()
            a.record()
            torch.autograd.grad(tv_out, f, gout)
            b.record(); b.synchronize()
            entry["torchvision_cuda_bwd_ms"] = a.elapsed_time(b)
            del f, tv_out
        except Exception as e:
            entry["torchvision_cuda"] = "unavailable: " + repr(e)[:120]
        out["roi_align_backward_c2"] = entry
        del gout, rws
        torch.cuda.empty_cache()
    d = synth.nms_stress_boxes(20000, synth.SEED_C5).to(dev)
    seg = torch.tensor([0, 20000], dtype=torch.int32, device=dev)
    boxes, scores = d[:, :4].contiguous(), d[:, 4].contiguous()
    ms = timed(lambda: ops.nms_batched(boxes, scores, seg, 0.7, 0, False), reps=5)
    out["nms_20k_boxes_c5"] = {"ms": ms, "pairs": 20000 * 19999 // 2, "giou_per_s": 20000 * 19999 / 2 / ms / 1e6,
                               "bound": "fp32 ALU + serial greedy chain (SURVEY 8d)"}
    d5 = synth.nms_stress_boxes(5000, synth.SEED_C5 + 1).to(dev)
    seg5 = torch.tensor([0, 5000], dtype=torch.int32, device=dev)
    ms = timed(lambda: ops.soft_nms_batched(d5, seg5, 0.5, 0.7, 0.1, 2), reps=5)
    out["soft_nms_5k_boxes"] = {"ms": ms, "bound": "serial selection chain"}
    return out


# ------------------------------------------------------------------------------------------ roofline helper
def roi_align_algorithmic_bytes(bxyxy, B, C, H, W):
    """min(sum_roi C*4*(floor(x2)-floor(x1)+2)*(floor(y2)-floor(y1)+2) clipped to the map,
    B*C*H*W*4) read + N*C*9*4 write."""
    import numpy as np
    r = bxyxy.cpu().numpy().astype(np.float64)
    x1 = np.clip(np.floor(r[:, 1]), 0, W - 1)
    x2 = np.clip(np.floor(r[:, 3]) + 1, 0, W - 1)
    y1 = np.clip(np.floor(r[:, 2]), 0, H - 1)
    y2 = np.clip(np.floor(r[:, 4]) + 1, 0, H - 1)
    win = np.maximum(x2 - x1 + 1, 0) * np.maximum(y2 - y1 + 1, 0)
    read = min(float(win.sum()) * C * 4, float(B) * C * H * W * 4)
    return read + r.shape[0] * C * 9 * 4, read, float(win.sum()) * C * 4


# ------------------------------------------------------------------------------------------ product arm
def run_product(args):
    import torch.distributed as dist
    from rrnet_b200 import ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ops._lib.lib()

    w = WORKLOAD
    B, C, H, W, K, Cf = w["B"], w["C"], w["H"], w["W"], w["K"], w["feat_ch"]
    seed = synth.SEED_C2 if world == 1 else synth.SEED_C4 + rank
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(synth.SEED_C2)
    host = {k: v.pin_memory() for k, v in x.items()}           # e2e inputs live in pinned host memory
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    folded = ops.head_fold({k: v.to(dev) for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, device=dev, roi_algo=args.roi_algo, head_algo=args.head_algo)
    gathered = None
    if world > 1:
        gathered = torch.empty(world * path.result_blob.numel(), dtype=torch.float32, device=dev)

    graph = None if args.no_graph else path.capture(d["hm"], d["wh"], d["off"], d["feat"])

    # Batches in flight.  At B = 8 a fifth of a step is spent in kernels that cannot fill the GPU (one CTA per image
    # or per class segment: top-K selection, greedy NMS scan, RoIAlign bookkeeping).  Steps are independent, so
    # consecutive steps alternate between `--streams` streams (own buffers, own CUDA graph) and the persistent
    # kernels leave `--sm-reserve` SMs free: the latency chains of one batch run next to the bandwidth-bound
    # kernels of the other.  Same launches, same results per step; `single_batch` in the line is one batch at a time.
    n_streams = 1 if graph is None else max(1, args.streams)
    pipes = []
    if n_streams > 1:
        ops.set_sm_reserve(args.sm_reserve)         # grid sizes are baked in at capture
        for _ in range(n_streams):
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                pp = ops.EvalPath(B, C, H, W, K, folded, device=dev, roi_algo=args.roi_algo, head_algo=args.head_algo)
                gg = pp.capture(d["hm"], d["wh"], d["off"], d["feat"])
            gb = torch.empty_like(gathered) if world > 1 else None
            pipes.append((st, pp, gg, gb))
        ops.set_sm_reserve(0)
        torch.cuda.synchronize()

    def step(events=None):
        if graph is not None and events is None:
            graph.replay()                 # the same launches, submitted as one CUDA graph
        else:
            path.forward(d["hm"], d["wh"], d["off"], d["feat"], stage_events=events)
        if world > 1:            # all-gather of detections for mAP (padded rows + per-image counts, one buffer)
            dist.all_gather_into_tensor(gathered, path.result_blob)

    def run_steps(n):
        """n steps on the current stream, or round-robin over the pipes (fork / join around them)."""
        if not pipes:
            for _ in range(n):
                step()
            return
        main_stream = torch.cuda.current_stream()
        for st, _, _, _ in pipes:
            st.wait_stream(main_stream)
        for i in range(n):
            st, pp, gg, gb = pipes[i % len(pipes)]
            with torch.cuda.stream(st):
                gg.replay()
                if world > 1:
                    dist.all_gather_into_tensor(gb, pp.result_blob)
        for st, _, _, _ in pipes:
            main_stream.wait_stream(st)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    launches0 = ops._lib.launch_count()
    path.forward(d["hm"], d["wh"], d["off"], d["feat"])
    launches_per_step = ops._lib.launch_count() - launches0     # kernels of ours per step (same inside the graph)
    for _ in range(max(args.warmup, 3)):
        step()
    run_steps(max(args.warmup, 3) * max(1, len(pipes)))
    sync_all()

    # ---- timed region: K steps, device-resident inputs, no host sync inside ----
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    t_beg.record()
    run_steps(args.steps)
    t_end.record()
    sync_all()
    clocks = sampler.stop()
    single = None
    if pipes:                      # one batch at a time (one stream, all SMs), for comparison
        s_beg, s_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_beg.record()
        for i in range(args.steps):
            step()
        s_end.record()
        sync_all()
        ts = torch.tensor([s_beg.elapsed_time(s_end)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        single = {"ms_per_step": float(ts.item()) / args.steps,
                  "value": world * B * args.steps / (float(ts.item()) / 1e3), "unit": UNIT}
        for _, pp, _, _ in pipes:  # every pipe computed what the single path computed
            if not (torch.equal(pp.s2, path.s2) and torch.equal(pp.counts, path.counts)):
                raise SystemExit("bench.py: a pipelined batch differs from the single-batch result")
    launches = launches_per_step * args.steps
    elapsed_ms = t_beg.elapsed_time(t_end)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms / 1e3)

    # ---- per-stage device times: a second, eager loop with events recorded by the library between stages ----
    n_ev = min(args.steps, 20)
    stage_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(n_ev)]
    for i in range(n_ev):
        step(stage_ev[i])
    sync_all()

    names = ["decode", "stage1_nms", "roi_align", "head", "generate_bbox"]
    stage_ms = {n: 0.0 for n in names}
    for evs in stage_ev:
        for j, n in enumerate(names):
            stage_ms[n] += evs[j].elapsed_time(evs[j + 1])
    stage_ms = {n: v / n_ev for n, v in stage_ms.items()}

    # ---- e2e: pinned host inputs -> H2D -> path -> D2H of the result, every step ----
    res_host = torch.empty(B * K, 6, dtype=torch.float32).pin_memory()
    cnt_host = torch.empty(B + 1, dtype=torch.int32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = res_host.numel() * 4 + cnt_host.numel() * 4

    # Double-buffered: the H2D copy of step i+1 (copy stream) overlaps the kernels and the D2H of step i
    # (compute stream); every step still moves all of its inputs from pinned host memory and reads its
    # result back, all inside the timed region.  PCIe carries 1.13 GB per step, so this leg is copy bound.
    d2 = {k: torch.empty_like(v) for k, v in d.items()}
    bufs = (d, d2)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(free[b])               # step i-2 no longer reads this buffer
                for k in ("hm", "wh", "off", "feat"):
                    bufs[b][k].copy_(host[k], non_blocking=True)
                h2d_done[b].record(copy_stream)
            main_stream.wait_event(h2d_done[b])
            path.forward(bufs[b]["hm"], bufs[b]["wh"], bufs[b]["off"], bufs[b]["feat"])
            if world > 1:
                dist.all_gather_into_tensor(gathered, path.result_blob)
            free[b].record(main_stream)
            res_host.copy_(path.s2, non_blocking=True)
            cnt_host.copy_(path.counts, non_blocking=True)

    e2e_run(2)
    sync_all()
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record()
    e2e_run(e2e_steps)
    e_end.record()
    sync_all()
    te = torch.tensor([e_beg.elapsed_time(e_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(te.item()) / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the step) ----
    r = path.results()
    peak, peak_src = measured_peaks()
    dominant = max(stage_ms, key=stage_ms.get)
    algo_total, algo_read, window_bytes = roi_align_algorithmic_bytes(r["bxyxy"], B, Cf, H, W)
    # fused eval path: the RoI feature tensor is never written (the head sums the partial slots), so the
    # RoIAlign stage's algorithmic bytes are the feature reads only ("0 write when fused with head")
    roof_bytes = {
        "decode": B * C * H * W * 4 + B * K * 16 + B * K * 32,
        "roi_align": algo_read,
    }
    traffic = {}
    tp = os.path.join(REPO, "profiles", "traffic.json")       # dram bytes per launch from the ncu --set full capture
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    roofline = None
    if dominant in roof_bytes:
        ach = roof_bytes[dominant] / (stage_ms[dominant] * 1e-3) / 1e9
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic.get(dominant), "algorithmic_bytes": roof_bytes[dominant],
                    "peak_source": peak_src}
    else:       # head: fp32 FFMA contraction, 49 valid 3x3 taps -> 993,280 flop per RoI
        tf = r["n"] * 993280.0 / (stage_ms[dominant] * 1e-3) / 1e12
        roofline = {"kernel": dominant, "bound": "tensor", "achieved": tf, "peak": 1624.9, "unit": "TFLOP/s",
                    "frac": tf / 1624.9, "traffic": None, "note": "fp32 FFMA kernel measured against the bf16 tensor peak"}
    roi_ach = algo_read / (stage_ms["roi_align"] * 1e-3) / 1e9
    dec_ach = roof_bytes["decode"] / (stage_ms["decode"] * 1e-3) / 1e9
    extra_roof = {"roi_align_gbs": roi_ach, "roi_align_frac": roi_ach / peak, "decode_gbs": dec_ach,
                  "decode_frac": dec_ach / peak, "roi_window_bytes_l2": window_bytes,
                  "head_tflops_fp32": r["n"] * 993280.0 / (stage_ms["head"] * 1e-3) / 1e12}

    # ---- the reference's own CUDA/torch path on the same batch, and the other scope rows, beside it ----
    ref_cuda = None
    aux = None
    if not args.no_aux:
        try:
            ms = cuda_reference(d, {k: v.to(dev) for k, v in hp.items()}, K)
            ref_cuda = {"ms_per_step": ms, "value": B / (ms / 1e3), "unit": UNIT,
                        "what": "reference call sequence (torch + torchvision CUDA ops, cuDNN head, fp32) on the same "
                                "device-resident batch, wall clock incl. its host syncs"}
        except Exception as e:                                       # torchvision CUDA ops missing on the box
            ref_cuda = {"unavailable": repr(e)[:200]}
        aux = aux_kernels(dev, peak, feat=d["feat"], rois=r["bxyxy"])

    # ---- CPU baseline beside it (rank 0, bounded sample) ----
    cpu_info = None
    if not args.no_cpu:
        _, cpu_info = cpu_reference(steps=args.cpu_steps, warmup=1)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, **WORKLOAD, "global_batch": world * B,
                       "parallelism": "image-sharded x%d, all-gather of detections" % world if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (features 1.07 GB per step)",
                       "submission": "eager launches" if args.no_graph else "CUDA graph replay of the step (memset + 14 kernels)",
                       "batches_in_flight": max(1, len(pipes)), "sm_reserve": args.sm_reserve if pipes else 0,
                       "rois_per_step": r["n"]},
            "clocks": clocks, "gpu_launches": int(launches) * world,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "single_batch": single, "roofline": roofline, "stages_ms": stage_ms, "kernels": extra_roof, "cpu_baseline": cpu_info,
            "reference_cuda": ref_cuda, "aux": aux}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-aux", action="store_true", help="skip the reference-CUDA leg and the aux kernel timings")
    ap.add_argument("--head-algo", type=int, default=0, help="0 tcgen05 tensor-core head (default), 1 fp32 FFMA head")
    ap.add_argument("--roi-algo", type=int, default=0, help="0 tile-centric RoIAlign with TMA-staged tiles (default), 1 direct gather, 4 tile-centric with load-staged tiles")
    ap.add_argument("--streams", type=int, default=2, help="batches in flight (steps alternate between that many streams)")
    ap.add_argument("--sm-reserve", type=int, default=8, help="SMs the persistent kernels leave free when --streams > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
