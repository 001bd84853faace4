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
NOMINAL_HBM_GBS = 8000.0            # BASELINE.json north_star: "B200 peak of about 8 TB/s"

# BASELINE.json configs[1..4] (SURVEY 8d "Configs restated"); --config picks one, the default is config 2
# (the configuration the metric is quoted on; with --gpus N it is run weak-scaled, B=8 per GPU).
CONFIGS = {
    2: dict(name=WORKLOAD_NAME, B=8, K=1500, seed="SEED_C2"),
    4: dict(name="rrnet_eval_image_sharded_b64_1088x1920_k1500", B=64, K=1500, seed="SEED_C4"),       # B split over the ranks
    5: dict(name="rrnet_eval_dense_scene_1088x1920_b16_k5000", B=16, K=5000, seed="SEED_C5"),
}


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    """Dense bf16 tensor TFLOP/s of this pool's B200s (MEASURED_PEAKS.json bf16_tflops), else the recipe's fallback."""
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("bf16_tflops", "bf16_tf", "tensor_bf16_tflops"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json %s)" % k
        except Exception:
            pass
    return 1600.0, "fallback (B200_PROFILING.md dense bf16)"


def roof_row(kernel, ms, nbytes, peak, bound="hbm", **extra):
    """One per-kernel roofline entry: achieved GB/s of the kernel's designed / algorithmic bytes against the measured
    copy bandwidth and against the nominal 8 TB/s (SURVEY 8d asks for both)."""
    gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    row = {"kernel": kernel, "bound": bound, "ms": ms, "bytes": int(nbytes), "achieved": gbs, "unit": "GB/s",
           "peak": peak, "frac": gbs / peak, "frac_nominal_8tbs": gbs / NOMINAL_HBM_GBS}
    row.update(extra)
    return row


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
def cpu_reference(steps, warmup, sample_images=1, seed=None, K=None):
    """Time the reference's CPU call sequence (oracle/ref_port.py) on `sample_images` images of the
    workload per step (config 2 unless K says otherwise), with every host thread torch can use.  Returns (img/s, info)."""
    from oracle import ref_port
    from rrnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = dict(WORKLOAD)
    if K is not None:
        w["K"] = K
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
            "sample": "%d image(s) of the workload per step (10x272x480 heat-map, K=%d, 256x272x480 "
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
    if args.config == 3:
        return run_train_reference(args)
    cfg = CONFIGS[args.config]
    if cfg["K"] > 1500:                        # a K=5000 image takes ~3x longer on the CPU: keep the arm within minutes
        steps = min(steps, 30)
    ips, info = cpu_reference(steps, warmup, K=cfg["K"])
    wl = dict(WORKLOAD, B=cfg["B"], K=cfg["K"])
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": info["ms_per_image"], "higher_is_better": True,
            "scaling": "strong" if args.config == 4 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "baseline_config": args.config,
                       "step": "1 image (bounded sample of the batch-%d step)" % cfg["B"], **wl},
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
    ms = timed(lambda: ops.focal_render_fwd_bwd(z, annos, n_obj, 512, 512))
    fb = 2 * n * 4 + annos.numel() * 4            # logits read once + gradient write
    out["focal_render_fused_fwd_bwd_c3"] = {"ms": ms, "algorithmic_bytes": fb, "gbs": fb / ms / 1e6,
                                            "frac_of_hbm_peak": fb / ms / 1e6 / peak,
                                            "unfused_pair_ms": out["render_targets_c3"]["ms"] + out["focal_fwd_bwd_c3"]["ms"],
                                            "replaces": "render_targets + focal_fwd_bwd (84 MB of traffic); one pass, target tiles in shared memory"}
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
    # stage-1 head tail at config 2: the heat-map head's final 1x1 conv (256 -> 10) + decode, with and
    # without the fusion.  Without: the reference's layer through cuDNN (fp32, TF32 off) writes the map, decode samples
    # and streams it again; with: rr_hm_tail_collect (one pass over t) and decode starts at the selection.
    if feat is not None:
        Bt, Ct, Ht, Wt = feat.shape
        Kt, Cc = 1500, 10
        t_hm = torch.relu(feat)                                        # stands in for the 3x3 conv + ReLU output
        wt = (torch.randn(Cc, Ct, 1, 1, generator=g) * (2.0 / Ct ** 0.5)).to(dev)
        bt = torch.full((Cc,), -2.19, device=dev)
        wh_t, off_t = [v.to(dev) for v in synth.wh_offset(Bt, Ht, Wt, synth.SEED_C2)]
        dws = torch.empty(ops._lib.lib().rr_decode_workspace_bytes(Bt, Cc, Ht, Wt, Kt), dtype=torch.uint8, device=dev)
        hm_buf = torch.empty(Bt, Cc, Ht, Wt, device=dev)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ms_conv = timed(lambda: torch.nn.functional.conv2d(t_hm, wt, bt), reps=5)
            hm_ref = torch.nn.functional.conv2d(t_hm, wt, bt)
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        ms_dec = timed(lambda: ops.decode_topk(hm_ref, wh_t, off_t, Kt, want_inds=False), reps=5)
        ms_tail = timed(lambda: ops.hm_tail_collect(t_hm, wt, bt, Kt, ws=dws, hm_out=hm_buf), reps=5)
        ms_sel = timed(lambda: ops.decode_topk(hm_buf, wh_t, off_t, Kt, want_inds=False, precollected_ws=dws), reps=5)
        tb = t_hm.numel() * 4 + hm_buf.numel() * 4
        out["hm_tail_fusion_c2"] = {
            "unfused_ms": ms_conv + ms_dec, "fused_ms": ms_tail + ms_sel,
            "unfused": {"conv1x1_cudnn_fp32_ms": ms_conv, "decode_sample_collect_select_ms": ms_dec},
            "fused": {"rr_hm_tail_collect_ms": ms_tail, "decode_select_only_ms": ms_sel},
            "algorithmic_bytes": tb, "gbs": tb / ms_tail / 1e6, "frac_of_hbm_peak": tb / ms_tail / 1e6 / peak,
            "frac_nominal_8tbs": tb / ms_tail / 1e6 / NOMINAL_HBM_GBS,
            "max_abs_logit_diff_vs_cudnn": float((hm_buf - hm_ref).abs().max()),
            "what": "t [8,256,272,480] read once + logits [8,10,272,480] written once (the sample pass adds ~4 %% of t)"}
        del t_hm, hm_ref, hm_buf, dws
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


def nms_baselines(dev):
    """The reference's own NMS implementations next to ours, same boxes, host arrays in / keep list out (wall clock,
    best of 3): `_nms` = the reference's CUDA kernel + host driver (ext/nms/nms/nms_kernel.cu:34-144, compiled unchanged
    for sm_100a into oracle/_ref/libref_gpu_nms.so) against rr_nms_legacy_host (same ABI); cpu_nms / cpu_soft_nms
    (ext/nms/nms/cpu_nms.pyx, oracle/_ref) against the device NMS / soft-NMS kernels.  Baseline leg: the only place
    besides cpu_baseline where oracle/_ref is executed."""
    import numpy as np
    from oracle import build_ref
    from rrnet_b200 import ops, synth
    ref_gpu = build_ref.load_gpu_nms()
    ref_cpu = build_ref.load()

    def best(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return 1e3 * min(ts), out

    rows = {}
    for n in (1500, 5000, 20000):
        dets = synth.nms_stress_boxes(n, synth.SEED_C5 + n).numpy()
        order = np.argsort(-dets[:, 4], kind="stable")
        ds = np.ascontiguousarray(dets[order])
        e = {"boxes": n}
        ms, keep = best(lambda: ops.nms_legacy_host(ds, 0.7))
        e["rr_nms_legacy_host_ms"] = ms
        e["kept"] = int(len(keep))
        if ref_gpu is not None:
            ms_r, keep_r = best(lambda: ref_gpu(ds, 0.7))
            e["reference_cuda_nms_ms"] = ms_r
            e["same_keep_list"] = bool(np.array_equal(np.asarray(keep), np.asarray(keep_r)))
            e["speedup_vs_reference_cuda"] = ms_r / ms
        if ref_cpu is not None:
            reps = 1 if n >= 20000 else 2
            ms_c, keep_c = best(lambda: ref_cpu.cpu_nms(ds, 0.7), reps=reps)
            e["reference_cpu_nms_ms"] = ms_c
            dd = torch.from_numpy(ds).to(dev)
            seg = torch.tensor([0, n], dtype=torch.int32, device=dev)
            bx, sc = dd[:, :4].contiguous(), dd[:, 4].contiguous()
            ms_d, _ = best(lambda: ops.nms_batched(bx, sc, seg, 0.7, 1, True))       # cpu_nms semantics: +1, >=
            e["rr_nms_batched_device_ms"] = ms_d
            if n <= 5000:
                ms_s, _ = best(lambda: ref_cpu.cpu_soft_nms(ds.copy(), 0.5, 0.7, 0.1, 2), reps=reps)
                e["reference_cpu_soft_nms_ms"] = ms_s
                ms_sd, _ = best(lambda: ops.soft_nms_batched(dd[:, :5].clone(), seg, 0.5, 0.7, 0.1, 2))
                e["rr_soft_nms_device_ms"] = ms_sd
        rows["n%d" % n] = e
    rows["note"] = ("wall clock ms, best of 3, one class; *_host/_nms/cpu_* take host arrays (their own H2D/D2H inside), "
                    "*_device_ms are the batched device kernels on resident boxes; cores=%d" % (os.cpu_count() or 1))
    return rows


# ------------------------------------------------------------------------------------------ config 3 (training loss path)
C3 = dict(B=32, C=10, h=128, w=128, img=512)
C3_NAME = "rrnet_train_loss_path_render_plus_focal_b32_512x512"
C3_METRIC = "images/sec Gaussian heat-map target render + focal loss (fwd+bwd) at 512x512"


def run_train_reference(args):
    """CPU arm of config 3: the oracle's C restatement of to_heatmap (datasets/transforms/functional.py:230-262) and of
    clamp(sigmoid) + focal_loss_for_hm forward and gradient (modules/loss/functional.py:25-51) on the box's cores."""
    import numpy as np
    import oracle
    from rrnet_b200 import synth
    oracle.build()
    cores = os.cpu_count() or 1
    oracle.set_threads(cores)
    B = 8                                           # bounded sample: 8 of the 32 images per step
    g = torch.Generator().manual_seed(synth.SEED_C3)
    z = (torch.randn(B, C3["C"], C3["h"], C3["w"], generator=g) * 2 - 2).numpy()
    annos = synth.train_annos(B, C3["img"], C3["img"], synth.SEED_C3)
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)

    def once():
        gt = np.stack([oracle.render(a.numpy(), C3["img"], C3["img"])["hm"] for a in annos])
        return oracle.focal(z, gt, want_grad=True)
    for _ in range(warmup):
        once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    dt = time.perf_counter() - t0
    ips = B * steps / dt
    info = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of the 32 images per step: oracle render (per image) + focal fwd+grad, %d steps" % (B, steps)}
    print(json.dumps({"impl": "reference", "metric": C3_METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": C3_NAME, "baseline_config": 3, **C3}, "cpu_baseline": info,
                      "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def run_train(args):
    """BASELINE.json configs[2]: Gaussian heat-map target render + focal loss (forward + backward), batch 32 at
    512x512 (logits / targets 32x10x128x128), one GPU per rank (weak scaling: the loss needs no collective, DDP averages
    the gradients).  A step = render the targets from the padded annotations, then the fused sigmoid+clamp+focal forward
    and backward; `fused` = the variant that never stores the target map."""
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
    B, C, h, w, img = C3["B"], C3["C"], C3["h"], C3["w"], C3["img"]
    g = torch.Generator().manual_seed(synth.SEED_C3 + rank)
    z_host = (torch.randn(B, C, h, w, generator=g) * 2 - 2).pin_memory()
    annos_h, nobj_h = synth.pad_annos(synth.train_annos(B, img, img, synth.SEED_C3 + rank))
    annos_h, nobj_h = annos_h.pin_memory(), nobj_h.pin_memory()
    z, annos, n_obj = z_host.to(dev), annos_h.to(dev), nobj_h.to(dev)
    n = z.numel()
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)       # 256 MB read between steps: cold L2
    sink = torch.zeros(1, dtype=torch.float32, device=dev)

    def one_step(zz, aa, nn):
        gt = ops.render_targets(aa, nn, img, img)[0]
        return ops.focal_fwd_bwd(zz, gt)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    l0 = ops._lib.launch_count()
    one_step(z, annos, n_obj)
    launches_per_step = ops._lib.launch_count() - l0
    for _ in range(max(args.warmup, 3)):
        one_step(z, annos, n_obj)
    # graph of (flush L2, step) and of (flush L2) alone: the difference is the step with a cold L2
    def graph_of(with_step):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            torch.sum(flush, dim=0, keepdim=True, out=sink)
            if with_step:
                one_step(z, annos, n_obj)
        return gph
    torch.cuda.synchronize()
    g_step, g_flush = graph_of(True), graph_of(False)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()

    def timed(gph):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        a.record()
        for _ in range(args.steps):
            gph.replay()
        b.record()
        sync_all()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_total = max(timed(g_step) - timed(g_flush), 1e-6)
    clocks = sampler.stop()
    value = world * B * args.steps / (ms_total / 1e3)

    # e2e: logits + annotations from pinned host memory every step, loss read back
    stats_host = torch.empty(4, dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 50))

    def e2e_run(k):
        for _ in range(k):
            zz, aa, nn = z_host.to(dev, non_blocking=True), annos_h.to(dev, non_blocking=True), nobj_h.to(dev, non_blocking=True)
            st, _ = one_step(zz, aa, nn)
            k = min(4, st.numel())
            stats_host[:k].copy_(st.reshape(-1)[:k], non_blocking=True)
    e2e_run(2)
    sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e2e_run(e2e_steps); b.record()
    sync_all()
    te = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(te.item()) / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    aux = aux_kernels(dev, peak)
    rows = []
    for key, name in (("render_targets_c3", "render_kernel (+ memset)"), ("focal_fwd_c3", "focal_forward_kernel"),
                      ("focal_fwd_bwd_c3", "focal_fwd_bwd_kernel"), ("focal_render_fused_fwd_bwd_c3", "focal_render_forward+backward (fused)"),
                      ("regl1_fwd_bwd_c3", "regl1_kernel (+ grad memset)")):
        e = aux[key]
        rows.append(roof_row(name, e["ms"], e["algorithmic_bytes"], peak, "hbm", what="SURVEY 8d algorithmic bytes, cold L2"))
    step_bytes = annos.numel() * 4 + n * 4 + 3 * n * 4            # render (annos in, map out) + focal fwd+bwd (3 maps)
    ach = step_bytes / (ms_total / args.steps * 1e-3) / 1e9
    roofline = {"kernel": "render + focal fwd+bwd (the step)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "frac_nominal_8tbs": ach / NOMINAL_HBM_GBS, "traffic": None,
                "algorithmic_bytes": step_bytes, "peak_source": peak_src}
    cpu_info = None
    if not args.no_cpu:
        import contextlib, io
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            run_train_reference(argparse.Namespace(steps=args.cpu_steps, warmup=1, gpus=1))
        cpu_info = json.loads(buf.getvalue())["cpu_baseline"]
    print(json.dumps({"metric": C3_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": C3_NAME, "baseline_config": 3, **C3, "global_batch": world * B,
                                 "l2": "256 MB read between steps (cold L2), its time subtracted",
                                 "submission": "CUDA graph replay", "objects": int(n_obj.sum())},
                      "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps * world),
                      "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 4 + annos.numel() * 4 + n_obj.numel() * 4),
                              "d2h_bytes_per_step": 16, "steps": e2e_steps},
                      "roofline": roofline, "rooflines": {"kernels": rows}, "cpu_baseline": cpu_info, "aux": aux}))
    if world > 1:
        dist.destroy_process_group()


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
def kernel_rooflines(path, d, peak, n_rois, algo_read, B, C, H, W, K, Cf, reps=5):
    """Per-kernel device times of one step (ops.KernelTrace: CUDA events between the library's launches, eager,
    queue kept full) and each kernel's roofline entry.  `bytes` is what the kernel moves BY DESIGN (formula in
    `what`); the stage-level SURVEY 8d algorithmic bytes are in the `stages` rows."""
    from rrnet_b200 import ops
    acc, order = {}, []
    for _ in range(reps):
        with ops.KernelTrace(capacity=64) as kt:
            path.forward(d["hm"], d["wh"], d["off"], d["feat"])
        for name, ms in kt.kernels:
            if name not in acc:
                acc[name] = 0.0
                order.append(name)
            acc[name] += ms / reps
    N = n_rois
    slots = 3.2 * N                                     # partial slots written (measured ~3.2 pieces per RoI at config 2)
    design = {
        "decode_sample_kernel": (B * 32768 * 4, "latency", "32 K sampled logits per image"),
        "decode_collect_kernel": (B * C * H * W * 4, "hbm", "heat-map logits read once: B*C*H*W*4"),
        "decode_select_kernel": (B * K * (8 + 16 + 32), "latency", "candidates + wh/offset gathers + rows out; one CTA per image"),
        "decode_select_cluster_kernel": (B * K * (2 * 8 + 16 + 32), "latency", "~2K candidates of 8 bytes per image + wh/offset gathers + rows out; a cluster of 8 CTAs per image"),
        "tail_sample_kernel": (0, "latency", "sampled pixel lines of the head activations"),
        "tail_thresh_kernel": (0, "latency", "threshold from 1024 folded sample maxima per image"),
        "tail_conv_collect_kernel": (B * Cf * H * W * 4 + B * C * H * W * 4, "hbm", "head activations read once + logits written once"),
        "stage1_partition_kernel": (B * K * (24 + 28), "latency", "rows in, class-sorted rows out; one CTA per image"),
        "nms_mask_kernel": (B * K * 20, "alu", "sorted boxes in; IoU tiles (fp32 ALU bound, SURVEY 8d)"),
        "nms_scan_kernel": (B * K * 8, "latency", "greedy scan per (image,class) segment: serial chain"),
        "stage1_compact_kernel": (B * K * 28 * 2, "latency", "kept rows in/out; one CTA per image"),
        "roi_prep_kernel": (N * (20 + 64 + 3 * 64 * 4 + 64 * 16), "latency", "RoI geometry + separable weight tables"),
        "roi_scan_kernel": (N * 8, "latency", "prefix sums over RoIs and tiles"),
        "roi_fill_kernel": (N * 3.2 * (32 + 384 + 384), "latency", "per-tile piece lists (descriptor + weight slices)"),
        "roi_tile_tma_kernel": (algo_read + slots * 9 * Cf * 4, "hbm", "feature map once (SURVEY 8d read bytes) + partial slots written"),
        "roi_tile_kernel": (algo_read + slots * 9 * Cf * 4, "hbm", "feature map once + partial slots written"),
        "roi_direct_kernel": (0, "latency", "RoIs routed to the direct path (none at config 2)"),
        "roi_combine_kernel": (slots * 9 * Cf * 4 + N * 9 * Cf * 4, "hbm", "partial slots in, RoI features out"),
        "head_tc_kernel": (slots * 9 * Cf * 4 + N * 16, "tensor", "partial slots in, [N,4] out; 3xTF32 tcgen05"),
        "head_forward_kernel": (N * 9 * Cf * 4 + N * 16, "alu", "fp32 FFMA head"),
        "generate_bbox_kernel": (N * 92, "latency", "N*(20+16+8) in, N*48 out"),
    }
    rows = []
    for name in order:
        nbytes, bound, what = design.get(name, (0, "latency", ""))
        extra = {"what": what}
        if name in ("head_tc_kernel", "head_forward_kernel"):
            tf = N * 993280.0 / (acc[name] * 1e-3) / 1e12
            tpeak, tsrc = measured_tensor_peak()
            extra.update({"useful_tflops_fp32": tf, "tensor_peak_bf16_tflops": tpeak, "tensor_peak_source": tsrc,
                          "frac_of_bf16_tensor_peak": tf / tpeak,
                          "note": "3xTF32 (3 MMAs per product) for 1e-5 fp32 parity; tensor-pipe-active %% is in profiles/"})
        rows.append(roof_row(name, acc[name], nbytes, peak, bound, **extra))
    return rows, sum(acc.values())


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
    ops.set_pdl(args.pdl)

    cfg = CONFIGS[args.config]
    w = WORKLOAD
    C, H, W, Cf = w["C"], w["H"], w["W"], w["feat_ch"]
    K = cfg["K"]
    if args.config == 4:                       # BASELINE.json configs[3]: batch 64 SPLIT over the GPUs, seeds 404 + rank
        if 64 % world:
            raise SystemExit("bench.py --config 4 needs 64 % gpus == 0")
        B, scaling = 64 // world, "strong"
        seed = synth.SEED_C4 + rank
    else:                                      # weak scaling: every rank runs the config's batch on its own images
        B, scaling = cfg["B"], "weak"
        seed = getattr(synth, cfg["seed"]) if world == 1 else synth.SEED_C4 + rank
    x = synth.eval_inputs(B, H, W, K, seed)
    hp = synth.head_params(synth.SEED_C2)
    host = {k: v.pin_memory() for k, v in x.items()}           # e2e inputs live in pinned host memory
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    folded = ops.head_fold({k: v.to(dev) for k, v in hp.items()})
    path = ops.EvalPath(B, C, H, W, K, folded, device=dev, roi_algo=args.roi_algo, head_algo=args.head_algo)

    # Exchange of the detections for mAP: every step's result blob (padded rows + per-image counts) goes
    # into a device ring of `--gather-every` slots and ONE NCCL all-gather moves the whole ring when it is full, so the
    # ranks rendezvous once per M steps instead of once per step; every step's detections are still exchanged inside the
    # timed region (a partly filled ring is flushed at the end of it).
    M = max(1, args.gather_every)
    blob_n = path.result_blob.numel()

    class Exchange:
        def __init__(self):
            self.ring = torch.empty(M, blob_n, dtype=torch.float32, device=dev) if world > 1 else None
            self.gathered = torch.empty(world, M, blob_n, dtype=torch.float32, device=dev) if world > 1 else None
            self.fill = 0
            self.collectives = 0

        def push(self, blob):
            if world == 1:
                return
            self.ring[self.fill].copy_(blob, non_blocking=True)
            self.fill += 1
            if self.fill == M:
                self.flush()

        def flush(self):
            if world > 1 and self.fill:
                dist.all_gather_into_tensor(self.gathered.view(-1), self.ring.view(-1))
                self.fill = 0
                self.collectives += 1

    xch = Exchange()
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
        # the cluster selection (8 CTAs x 144 KB per image) cannot be placed while the other batch's persistent kernels
        # hold 140 SMs: with two batches in flight the one-CTA-per-image selection fits into the reserved SMs instead
        ops.set_option(ops.OPT_SELECT_SINGLE_CTA, 1)
        for _ in range(n_streams):
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                pp = ops.EvalPath(B, C, H, W, K, folded, device=dev, roi_algo=args.roi_algo, head_algo=args.head_algo)
                gg = pp.capture(d["hm"], d["wh"], d["off"], d["feat"])
                xx = Exchange()
            pipes.append((st, pp, gg, xx))
        ops.set_sm_reserve(0)
        ops.set_option(ops.OPT_SELECT_SINGLE_CTA, 0)
        torch.cuda.synchronize()

    def step(events=None):
        if graph is not None and events is None:
            graph.replay()                 # the same launches, submitted as one CUDA graph
        else:
            path.forward(d["hm"], d["wh"], d["off"], d["feat"], stage_events=events)
        xch.push(path.result_blob)

    def run_steps(n):
        """n steps on the current stream, or round-robin over the pipes (fork / join around them)."""
        if not pipes:
            for _ in range(n):
                step()
            xch.flush()
            return
        main_stream = torch.cuda.current_stream()
        for st, _, _, _ in pipes:
            st.wait_stream(main_stream)
        for i in range(n):
            st, pp, gg, xx = pipes[i % len(pipes)]
            with torch.cuda.stream(st):
                gg.replay()
                xx.push(pp.result_blob)
        for st, _, _, xx in pipes:
            with torch.cuda.stream(st):
                xx.flush()
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
    xch.flush()
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
        xch.flush()
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
    if world > 1:                  # the exchange delivered every rank's rows: rank r's slot holds rank r's counts
        src = pipes[0] if pipes else None
        g_all = (src[3] if src else xch).gathered
        mine = (src[1] if src else path).counts
        if not torch.equal(g_all[rank, M - 1, B * K * 6:].view(torch.int32), mine):
            raise SystemExit("bench.py: all-gathered detections do not contain this rank's counts")
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
    xch.flush()
    sync_all()

    names = ["decode", "stage1_nms", "roi_align", "head", "generate_bbox"]
    stage_ms = {n: 0.0 for n in names}
    for evs in stage_ev:
        for j, n in enumerate(names):
            stage_ms[n] += evs[j].elapsed_time(evs[j + 1])
    stage_ms = {n: v / n_ev for n, v in stage_ms.items()}

    # ---- variant: the features arrive already ReLU-ed.  In the model they do: forward_stage1 builds relu(pre_feat[-1]) for
    # the stage-1 heads (models/rrnet.py:144) and the host mirror hands that tensor to the path (EvalPath feat_is_relu), so
    # RoIAlign skips its own ReLU.  The headline numbers above keep the fused ReLU (N(0,1) features). ----
    variant = None
    if not args.no_aux and world == 1:
        try:
            feat_r = torch.relu(d["feat"])
            pv = ops.EvalPath(B, C, H, W, K, folded, device=dev, roi_algo=args.roi_algo, head_algo=args.head_algo, feat_is_relu=True)
            gv = pv.capture(d["hm"], d["wh"], d["off"], feat_r)
            v_beg, v_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            v_beg.record()
            for _ in range(args.steps):
                gv.replay()
            v_end.record()
            torch.cuda.synchronize()
            v_ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            pv.forward(d["hm"], d["wh"], d["off"], feat_r, stage_events=v_ev)
            torch.cuda.synchronize()
            same = bool(torch.equal(pv.s2, path.s2) and torch.equal(pv.counts, path.counts))
            ms_v = v_beg.elapsed_time(v_end) / args.steps
            variant = {"feat_is_relu": {"single_batch_ms_per_step": ms_v, "value": B / (ms_v / 1e3), "unit": UNIT,
                                        "roi_align_stage_ms": v_ev[2].elapsed_time(v_ev[3]), "same_result": same}}
            del feat_r, pv, gv
            torch.cuda.empty_cache()
        except Exception as e:
            variant = {"feat_is_relu": {"error": repr(e)[:200]}}

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
            xch.push(path.result_blob)
            free[b].record(main_stream)
            res_host.copy_(path.s2, non_blocking=True)
            cnt_host.copy_(path.counts, non_blocking=True)
        xch.flush()

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

    if rank != 0:                  # the CPU legs run on rank 0 only; the other ranks are done (no spin barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: the dominant stage (the contract's object) + one entry per stage and per kernel ----
    r = path.results()
    peak, peak_src = measured_peaks()
    dominant = max(stage_ms, key=stage_ms.get)
    algo_total, algo_read, window_bytes = roi_align_algorithmic_bytes(r["bxyxy"], B, Cf, H, W)
    # fused eval path: the RoI feature tensor is never written (the head sums the partial slots), so the
    # RoIAlign stage's algorithmic bytes are the feature reads only ("0 write when fused with head")
    roof_bytes = {
        "decode": B * C * H * W * 4 + B * K * 16 + B * K * 32,
        "stage1_nms": B * K * 24 + r["n"] * 28,
        "roi_align": algo_read,
        "head": r["n"] * 9216 + r["n"] * 16 + 282640,
        "generate_bbox": r["n"] * 92,
    }
    bounds = {"decode": "hbm", "stage1_nms": "alu", "roi_align": "hbm", "head": "tensor", "generate_bbox": "hbm"}
    traffic = {}
    tp = os.path.join(REPO, "profiles", "traffic.json")       # dram bytes per launch from the ncu --set full capture
    if os.path.exists(tp) and args.config == 2 and B == 8:    # the capture is of the config-2 step
        traffic = json.load(open(tp))
    stage_rows = [roof_row(n, stage_ms[n], roof_bytes[n], peak, bounds[n], traffic=traffic.get(n),
                           what="stage, SURVEY 8d algorithmic bytes") for n in names]
    tpeak, tsrc = measured_tensor_peak()
    if dominant != "head":
        ach = roof_bytes[dominant] / (stage_ms[dominant] * 1e-3) / 1e9
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "frac_nominal_8tbs": ach / NOMINAL_HBM_GBS, "traffic": traffic.get(dominant),
                    "algorithmic_bytes": roof_bytes[dominant], "peak_source": peak_src}
    else:       # head: 49 valid 3x3 taps -> 993,280 flop per RoI, against the measured dense bf16 tensor peak
        tf = r["n"] * 993280.0 / (stage_ms[dominant] * 1e-3) / 1e12
        roofline = {"kernel": dominant, "bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": tf / tpeak, "traffic": traffic.get("head"), "peak_source": tsrc,
                    "note": "useful fp32 flops of a 3xTF32 kernel against the bf16 tensor peak"}
    kernel_rows, kernel_sum = kernel_rooflines(path, d, peak, r["n"], algo_read, B, C, H, W, K, Cf)
    roi_ach = algo_read / (stage_ms["roi_align"] * 1e-3) / 1e9
    dec_ach = roof_bytes["decode"] / (stage_ms["decode"] * 1e-3) / 1e9
    extra_roof = {"roi_align_gbs": roi_ach, "roi_align_frac": roi_ach / peak, "decode_gbs": dec_ach,
                  "decode_frac": dec_ach / peak, "roi_window_bytes_l2": window_bytes,
                  "head_tflops_fp32": r["n"] * 993280.0 / (stage_ms["head"] * 1e-3) / 1e12,
                  "sum_of_kernel_ms": kernel_sum}

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
    nms_base = None
    if not args.no_cpu:
        _, cpu_info = cpu_reference(steps=args.cpu_steps, warmup=1, K=K)
        if not args.no_aux:
            nms_base = nms_baselines(dev)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "baseline_config": args.config, "B": B, "C": C, "H": H, "W": W, "K": K,
                       "feat_ch": Cf, "global_batch": world * B,
                       "parallelism": ("image-sharded x%d, detections all-gathered (NCCL) every %d steps" % (world, M))
                       if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (features %.2f GB per step)" % (B * Cf * H * W * 4 / 1e9),
                       "submission": "eager launches" if args.no_graph else "CUDA graph replay of the step (memset + %d kernels)" % launches_per_step,
                       "batches_in_flight": max(1, len(pipes)), "sm_reserve": args.sm_reserve if pipes else 0,
                       "programmatic_dependent_launch": bool(args.pdl),
                       "rois_per_step": r["n"]},
            "clocks": clocks, "gpu_launches": int(launches) * world,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "single_batch": single, "roofline": roofline, "rooflines": {"stages": stage_rows, "kernels": kernel_rows},
            "stages_ms": stage_ms, "kernels": extra_roof, "cpu_baseline": cpu_info,
            "reference_cuda": ref_cuda, "aux": aux, "nms_baselines": nms_base, "variants": variant}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configuration: 2 eval path B=8 K=1500 (default; weak-scaled with --gpus), 3 training loss "
                         "path B=32 at 512x512, 4 eval batch 64 SPLIT over the GPUs (strong scaling), 5 dense scene B=16 K=5000")
    ap.add_argument("--pdl", action="store_true", help="programmatic dependent launch between the kernels of a step (measured: no gain for one batch at a time, slower with two in flight)")
    ap.add_argument("--gather-every", type=int, default=8,
                    help="multi-GPU: all-gather the detections of that many steps with one collective (device ring buffer)")
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
    elif args.config == 3:
        run_train(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
