"""Torch-tensor front end of the C ABI (include/rrnet_b200.h).

PyTorch is used here for device memory and streams only: every function validates its
tensors, allocates outputs / scratch with torch.empty on the current device and passes raw
pointers plus the current CUDA stream to librrnet_b200.so.  Nothing here computes on the CPU and
nothing falls back to torch ops: without the built library or without a CUDA device these
functions raise.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import RRNetB200Error, check

MAX_TOPK = 16384


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RRNetB200Error("%s must be a CUDA tensor (rrnet_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise RRNetB200Error("%s must be float32, got %s" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise RRNetB200Error("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    return t.contiguous()


def _i32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.int32:
        raise RRNetB200Error("%s must be a CUDA int32 tensor" % name)
    return t.contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------- decode
DECODE_RAW_SCORES = 0x100


PRECOLLECTED = 0x200          # RR_DECODE_PRECOLLECTED


def decode_topk(hm, wh, off, K, pool=0, want_inds=True, raw_scores=False, precollected_ws=None):
    """models/rrnet.py:117-138 on logits.  -> dets [B,K,6] (x1,y1,x2,y2,score,cls), inds [B,K] int64.
    raw_scores=True: hm already holds scores (RRNet._topk alone, :93-109); wh/off may then be None.
    precollected_ws: the workspace hm_tail_collect filled for this hm (skips the sample + collect passes)."""
    hm = _f32(hm, "hm", 4)
    B, C, H, W = hm.shape
    if raw_scores and wh is None and off is None:
        pool = int(pool) | DECODE_RAW_SCORES
    else:
        wh, off = _f32(wh, "wh", 4), _f32(off, "off", 4)
        if tuple(wh.shape) != (B, 2, H, W) or tuple(off.shape) != (B, 2, H, W):
            raise RRNetB200Error("wh/off must be [B,2,H,W] matching hm")
        if raw_scores:
            pool = int(pool) | DECODE_RAW_SCORES
    L = _lib.lib()
    dets = torch.empty(B, K, 6, dtype=torch.float32, device=hm.device)
    inds = torch.empty(B, K, dtype=torch.int64, device=hm.device) if want_inds else None
    if precollected_ws is not None:
        ws, pool = precollected_ws, int(pool) | PRECOLLECTED
    else:
        ws = _ws(L.rr_decode_workspace_bytes(B, C, H, W, K), hm.device)
    check(L.rr_decode_topk(_ptr(hm), _ptr(wh), _ptr(off), B, C, H, W, int(K), int(pool), _ptr(dets), _ptr(inds),
                           _ptr(ws), ws.numel(), _stream()), "rr_decode_topk")
    return dets, inds


# --------------------------------------------------------------------------------- NMS


def hm_tail_collect(t, weight, bias, K, ws=None, hm_out=None):
    """Last layer of the heat-map head (1x1 conv 256 -> C + bias, detectors/centernet_detector.py:11-15) fused with
    the decode's candidate collection: t [B,Cin,H,W] (after the head's 3x3 conv + ReLU), weight [C,Cin] or [C,Cin,1,1],
    bias [C] -> (hm logits [B,C,H,W], ws).  `ws` (a decode workspace, or EvalPath.ws) then holds the candidate lists:
    pass it with precollected=True to decode_topk / EvalPath.forward."""
    t = _f32(t, "t", 4)
    B, Cin, H, W = t.shape
    weight = _f32(weight.reshape(weight.shape[0], -1), "weight", 2)
    bias = _f32(bias, "bias", 1)
    C = weight.shape[0]
    if weight.shape[1] != Cin or bias.numel() != C:
        raise RRNetB200Error("hm_tail_collect: weight %s / bias %s do not match t %s" % (tuple(weight.shape), tuple(bias.shape), tuple(t.shape)))
    L = _lib.lib()
    if ws is None:
        ws = _ws(L.rr_decode_workspace_bytes(B, C, H, W, K), t.device)
    if hm_out is None:
        hm_out = torch.empty(B, C, H, W, dtype=torch.float32, device=t.device)
    check(L.rr_hm_tail_collect(_ptr(t), _ptr(weight), _ptr(bias), B, Cin, C, H, W, int(K), _ptr(hm_out), _ptr(ws),
                               ws.numel(), _stream()), "rr_hm_tail_collect")
    return hm_out, ws


def stage1_nms(dets, num_classes, thr=0.7):
    """RRNet.nms + the per-image loop of RRNet.forward for the batch, no host sync.
    -> bxyxy [B*K,5], scores [B*K], clses [B*K] (capacity rows), counts [B+1] int32 (device)."""
    dets = _f32(dets, "dets", 3)
    B, K, six = dets.shape
    if six != 6:
        raise RRNetB200Error("dets must be [B,K,6]")
    L = _lib.lib()
    dev = dets.device
    bxyxy = torch.empty(B * K, 5, dtype=torch.float32, device=dev)
    scores = torch.empty(B * K, dtype=torch.float32, device=dev)
    clses = torch.empty(B * K, dtype=torch.float32, device=dev)
    counts = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ws = _ws(L.rr_stage1_nms_workspace_bytes(B, K, num_classes), dev)
    check(L.rr_stage1_nms(_ptr(dets), B, K, int(num_classes), float(thr), _ptr(bxyxy), _ptr(scores), _ptr(clses),
                          _ptr(counts), _ptr(ws), ws.numel(), _stream()), "rr_stage1_nms")
    return bxyxy, scores, clses, counts


def nms_batched(boxes, scores, seg_offsets, thr, pixel_offset=0, ge_cmp=False):
    """Segmented hard NMS.  -> keep_idx [M] int32 (global rows, per segment at its offset), keep_count [S]."""
    boxes, scores = _f32(boxes, "boxes", 2), _f32(scores, "scores", 1)
    seg_offsets = _i32(seg_offsets, "seg_offsets")
    M, S = boxes.shape[0], seg_offsets.numel() - 1
    if boxes.shape[1] != 4 or scores.shape[0] != M or S < 1:
        raise RRNetB200Error("boxes [M,4], scores [M], seg_offsets [S+1] expected")
    L = _lib.lib()
    dev = boxes.device
    keep_idx = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    keep_cnt = torch.empty(S, dtype=torch.int32, device=dev)
    ws = _ws(L.rr_nms_workspace_bytes(M, S), dev)
    check(L.rr_nms_batched(_ptr(boxes), _ptr(scores), _ptr(seg_offsets), M, S, float(thr), int(pixel_offset),
                           int(bool(ge_cmp)), _ptr(keep_idx), _ptr(keep_cnt), _ptr(ws), ws.numel(), _stream()),
          "rr_nms_batched")
    return keep_idx, keep_cnt


def nms(boxes, scores, thr, pixel_offset=0, ge_cmp=False):
    """Single-set hard NMS -> kept row indices (int64, acceptance order).  One host sync for the count."""
    M = boxes.shape[0]
    if M == 0:
        return torch.empty(0, dtype=torch.int64, device=boxes.device)
    seg = torch.tensor([0, M], dtype=torch.int32, device=boxes.device)
    keep_idx, keep_cnt = nms_batched(boxes, scores, seg, thr, pixel_offset, ge_cmp)
    return keep_idx[: int(keep_cnt.item())].long()


def nms_legacy_host(dets_sorted, thresh, device_id=0):
    """The reference's `_nms` ABI (ext/nms/nms/gpu_nms.hpp): host rows sorted by score -> kept positions."""
    d = np.ascontiguousarray(dets_sorted, dtype=np.float32)
    n, dim = d.shape
    keep = np.zeros(max(n, 1), dtype=np.int32)
    num = ctypes.c_int(0)
    L = _lib.lib()
    check(L.rr_nms_legacy_host(keep.ctypes.data_as(ctypes.c_void_p), ctypes.cast(ctypes.byref(num), ctypes.c_void_p),
                               d.ctypes.data_as(ctypes.c_void_p), n, dim, float(thresh), int(device_id)),
          "rr_nms_legacy_host")
    return keep[: num.value]


def soft_nms_batched(boxes5, seg_offsets, sigma=0.5, Nt=0.3, threshold=0.001, method=1):
    """Soft-NMS per segment, on a COPY of boxes5 [M,5].  -> rows [M,5] (first keep_count[s] rows of each
    segment are the survivors, selection order, decayed scores), src_idx [M] int32, keep_count [S] int32."""
    boxes5 = _f32(boxes5, "boxes5", 2).clone()
    seg_offsets = _i32(seg_offsets, "seg_offsets")
    M, S = boxes5.shape[0], seg_offsets.numel() - 1
    if boxes5.shape[1] != 5 or S < 1:
        raise RRNetB200Error("boxes5 [M,5], seg_offsets [S+1] expected")
    L = _lib.lib()
    dev = boxes5.device
    src = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    cnt = torch.empty(S, dtype=torch.int32, device=dev)
    ws = _ws(L.rr_soft_nms_workspace_bytes(M), dev)
    check(L.rr_soft_nms_batched(_ptr(boxes5), _ptr(seg_offsets), M, S, float(sigma), float(Nt), float(threshold),
                                int(method), _ptr(src), _ptr(cnt), _ptr(ws), ws.numel(), _stream()),
          "rr_soft_nms_batched")
    return boxes5, src, cnt


# --------------------------------------------------------------------------------- RoIAlign / head / bbox
def roi_align(feat, rois, n_dev=None, relu=True, algo=0):
    """torchvision.ops.roi_align(relu(feat), rois, (3,3)) (models/rrnet.py:51) -> [n,C,3,3].
    algo 0 = tile-centric kernel (default; tiles arrive by TMA when W % 4 == 0), 1 = direct per-RoI gather,
    2 = tile-centric kernel with tiles staged by ordinary loads."""
    feat, rois = _f32(feat, "feat", 4), _f32(rois, "rois", 2)
    B, C, H, W = feat.shape
    n = rois.shape[0]
    if rois.shape[1] != 5:
        raise RRNetB200Error("rois must be [n,5]")
    out = torch.empty(n, C, 3, 3, dtype=torch.float32, device=feat.device)
    if n_dev is not None:
        n_dev = _i32(n_dev, "n_dev")
    if n == 0:
        return out
    L = _lib.lib()
    ws = _ws(L.rr_roi_align_workspace_bytes(n, B, C, H, W), feat.device)
    check(L.rr_roi_align(_ptr(feat), _ptr(rois), _ptr(n_dev), n, B, C, H, W, int(bool(relu)), int(algo), _ptr(out),
                         _ptr(ws), ws.numel(), _stream()), "rr_roi_align")
    return out


HEAD_KEYS = ("w1", "bn1", "w2", "bn2", "w3", "bn3", "wr", "br")
_HEAD_SHAPES = {"w1": (64, 256), "bn1": (4, 64), "w2": (64, 64, 3, 3), "bn2": (4, 64),
                "w3": (256, 64), "bn3": (4, 256), "wr": (4, 256), "br": (4,)}


def head_fold(params):
    """Fold the eval-mode BatchNorms into the convolutions once -> opaque folded parameter block."""
    L = _lib.lib()
    ts = []
    for k in HEAD_KEYS:
        t = _f32(params[k], k)
        if tuple(t.shape) != _HEAD_SHAPES[k]:
            raise RRNetB200Error("head param %s must have shape %s, got %s" % (k, _HEAD_SHAPES[k], tuple(t.shape)))
        ts.append(t)
    folded = torch.empty(L.rr_head_folded_floats(), dtype=torch.float32, device=ts[0].device)
    check(L.rr_head_fold(*[_ptr(t) for t in ts], _ptr(folded), _stream()), "rr_head_fold")
    return folded


def head_params_from_module(head_detector):
    """Collect HEAD_KEYS from a module with the reference's layout (detectors/fasterrcnn_detector.py:9-11)."""
    tl = head_detector.top_layer

    def bn(m):
        return torch.stack([m.weight, m.bias, m.running_mean, m.running_var]).detach().float().contiguous()

    return {
        "w1": tl.conv1.weight.detach().float().reshape(64, 256).contiguous(), "bn1": bn(tl.bn1),
        "w2": tl.conv2.weight.detach().float().contiguous(), "bn2": bn(tl.bn2),
        "w3": tl.conv3.weight.detach().float().reshape(256, 64).contiguous(), "bn3": bn(tl.bn3),
        "wr": head_detector.regressor.weight.detach().float().reshape(4, 256).contiguous(),
        "br": head_detector.regressor.bias.detach().float().contiguous(),
    }


def roi_align_backward(feat, rois, grad_out, relu=True, algo=0, n_dev=None, ws=None):
    """d loss / d feat of roi_align(relu(feat), rois, (3,3)): feat [B,C,H,W], rois [n,5], grad_out [n,C,3,3]
    -> grad_feat [B,C,H,W] (a gather on the tile path; atomics only for direct-path RoIs)."""
    feat = _f32(feat, "feat", 4)
    rois = _f32(rois, "rois", 2)
    grad_out = _f32(grad_out, "grad_out", 4)
    B, C, H, W = feat.shape
    n = rois.shape[0]
    if grad_out.shape != (n, C, 3, 3) or (n and rois.shape[1] != 5):
        raise RRNetB200Error("roi_align_backward: grad_out must be [n,C,3,3], rois [n,5]")
    grad = torch.empty_like(feat)
    L = _lib.lib()
    if ws is None:
        ws = _ws(L.rr_roi_align_workspace_bytes(max(n, 1), B, C, H, W), feat.device)
    if n_dev is not None:
        n_dev = _i32(n_dev, "n_dev")
    check(L.rr_roi_align_backward(_ptr(feat), _ptr(rois), _ptr(n_dev), n, B, C, H, W, int(bool(relu)), int(algo),
                                  _ptr(grad_out), _ptr(grad), _ptr(ws), ws.numel(), _stream()), "rr_roi_align_backward")
    return grad


def head_forward(roi_feat, folded, n_dev=None, algo=0):
    """Re-regression head on [n,256,3,3] RoI features -> [n,4].  algo 0 = tcgen05 tensor cores (3xTF32),
    1 = fp32 FFMA."""
    roi_feat = _f32(roi_feat, "roi_feat", 4)
    n = roi_feat.shape[0]
    if tuple(roi_feat.shape[1:]) != (256, 3, 3):
        raise RRNetB200Error("roi_feat must be [n,256,3,3]")
    reg = torch.empty(n, 4, dtype=torch.float32, device=roi_feat.device)
    if n_dev is not None:
        n_dev = _i32(n_dev, "n_dev")
    if n == 0:
        return reg
    check(_lib.lib().rr_head_forward(_ptr(roi_feat), _ptr(n_dev), n, _ptr(_f32(folded, "folded")), int(algo), _ptr(reg),
                                     _stream()), "rr_head_forward")
    return reg


def generate_bbox(bxyxy, reg, scores, clses, n_dev=None, scale=4.0):
    """RRNetOperator.generate_bbox for all rows -> s1 [n,6], s2 [n,6]."""
    bxyxy, reg = _f32(bxyxy, "bxyxy", 2), _f32(reg, "reg", 2)
    scores, clses = _f32(scores, "scores", 1), _f32(clses, "clses", 1)
    n = bxyxy.shape[0]
    s1 = torch.empty(n, 6, dtype=torch.float32, device=bxyxy.device)
    s2 = torch.empty(n, 6, dtype=torch.float32, device=bxyxy.device)
    if n_dev is not None:
        n_dev = _i32(n_dev, "n_dev")
    check(_lib.lib().rr_generate_bbox(_ptr(bxyxy), _ptr(reg), _ptr(scores), _ptr(clses), _ptr(n_dev), n, float(scale),
                                      _ptr(s1), _ptr(s2), _stream()), "rr_generate_bbox")
    return s1, s2


class EvalPath:
    """Pre-allocated, sync-free eval path for a fixed (B,C,H,W,K): decode -> stage-1 NMS ->
    RoIAlign+ReLU -> head -> generate_bbox in one C-ABI call (rr_eval_forward).  CUDA-graph capturable."""

    def __init__(self, B, C, H, W, K, head_folded, feat_ch=256, pool=0, nms_thr=0.7, scale=4.0, device=None,
                 keep_roi_feat=False, roi_algo=0, head_algo=0, feat_is_relu=False):
        L = _lib.lib()
        dev = torch.device(device if device is not None else "cuda")
        self.shape = (B, C, H, W, K, feat_ch)
        self.pool, self.nms_thr, self.scale = int(pool), float(nms_thr), float(scale)
        # bit 0: direct RoIAlign, bit 2 (roi_algo=4): load-staged tiles, bit 1: FFMA head, bit 3: feat is already relu(feat)
        self.roi_algo = int(roi_algo) | (int(head_algo) << 1) | (8 if feat_is_relu else 0)
        self.folded = _f32(head_folded, "head_folded")
        n = B * K
        f32 = dict(dtype=torch.float32, device=dev)
        self.dets = torch.empty(B, K, 6, **f32)
        self.inds = torch.empty(B, K, dtype=torch.int64, device=dev)
        self.bxyxy = torch.empty(n, 5, **f32)
        self.scores = torch.empty(n, **f32)
        self.clses = torch.empty(n, **f32)
        self.reg = torch.empty(n, 4, **f32)
        self.s1 = torch.empty(n, 6, **f32)
        # the final rows and the per-image counts share one buffer: a multi-GPU run exchanges them with ONE all-gather
        self.result_blob = torch.zeros(n * 6 + B + 1, **f32)
        self.s2 = self.result_blob[: n * 6].view(n, 6)
        self.counts = self.result_blob[n * 6:].view(torch.int32)
        self.roi_feat = torch.empty(n, feat_ch, 3, 3, **f32) if keep_roi_feat else None
        self.ws = _ws(L.rr_eval_workspace_bytes(B, C, H, W, K, feat_ch), dev)

    def forward_from_tail(self, t_hm, hm_weight, hm_bias, wh, off, feat, hm_out=None, stage_events=None):
        """The same path with the heat-map head's last layer fused in: t_hm [B,Cin,H,W] is the output of
        the head's 3x3 conv + ReLU; the 1x1 conv, the logit map and the candidate lists come from ONE pass over it
        (rr_hm_tail_collect), then rr_eval_forward starts at the top-K selection.  -> the logits [B,C,H,W]."""
        B, C, H, W, K, Cf = self.shape
        if hm_out is None:
            if getattr(self, "hm_buf", None) is None:
                self.hm_buf = torch.empty(B, C, H, W, dtype=torch.float32, device=self.ws.device)
            hm_out = self.hm_buf
        hm_tail_collect(t_hm, hm_weight, hm_bias, K, ws=self.ws, hm_out=hm_out)
        self.forward(hm_out, wh, off, feat, stage_events=stage_events, precollected=True)
        return hm_out

    def forward(self, hm, wh, off, feat, stage_events=None, precollected=False):
        """stage_events: optional list of 6 torch.cuda.Event(enable_timing=True) recorded by the library
        before decode and after each of decode / NMS / RoIAlign / head / bbox."""
        B, C, H, W, K, Cf = self.shape
        ev = None
        if stage_events is not None:
            for e in stage_events:
                e.record()                      # torch creates the CUDA event lazily on first record
            ev = (ctypes.c_void_p * 6)(*[e.cuda_event for e in stage_events])
        hm, wh, off, feat = _f32(hm, "hm", 4), _f32(wh, "wh", 4), _f32(off, "off", 4), _f32(feat, "feat", 4)
        if tuple(hm.shape) != (B, C, H, W) or tuple(feat.shape) != (B, Cf, H, W):
            raise RRNetB200Error("EvalPath was built for hm %s / feat %s" % ((B, C, H, W), (B, Cf, H, W)))
        check(_lib.lib().rr_eval_forward(
            _ptr(hm), _ptr(wh), _ptr(off), _ptr(feat), B, C, H, W, K, Cf, (PRECOLLECTED if precollected else self.pool), self.nms_thr,
            self.roi_algo, _ptr(self.folded), self.scale, _ptr(self.dets), _ptr(self.inds), _ptr(self.bxyxy), _ptr(self.scores),
            _ptr(self.clses), _ptr(self.counts), _ptr(self.reg), _ptr(self.s1), _ptr(self.s2), _ptr(self.roi_feat),
            _ptr(self.ws), self.ws.numel(), _stream(), ev), "rr_eval_forward")
        return self

    def capture(self, hm, wh, off, feat):
        """Capture forward() on these (static) input tensors into a CUDA graph; graph.replay() then re-runs
        the whole path (a memset + 14 kernels) with one launch.  The C entry points never allocate or
        synchronise, so they are capturable as they are."""
        self.forward(hm, wh, off, feat)                    # warm-up: function attributes, lazy module load
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward(hm, wh, off, feat)
        return g

    def results(self):
        """One host sync: -> dict of tensors sliced to the live row count + per-image counts (python list)."""
        counts = self.counts.tolist()
        n = counts[-1]
        return {"n": n, "counts": counts[:-1], "bxyxy": self.bxyxy[:n], "scores": self.scores[:n],
                "clses": self.clses[:n], "reg": self.reg[:n], "s1": self.s1[:n], "s2": self.s2[:n]}


def regl1_fwd_bwd(output, mask, ind, target, grad_scale=1.0, want_grad=True):
    """RegL1Loss (modules/loss/regl1loss.py:9-17) of a regression map and its gradient in one launch.
    output [B,c,H,W], mask [B,max_n(,1)], ind [B,max_n(,1)] (float or int), target [B,max_n,c]
    -> (loss [1], grad [B,c,H,W] * grad_scale or None)."""
    output = _f32(output, "output", 4)
    B, c, H, W = output.shape
    target = _f32(target, "target", 3)
    max_n = target.shape[1]
    mask = _f32(mask.reshape(B, max_n).float(), "mask", 2)
    ind = _f32(ind.reshape(B, max_n).float(), "ind", 2)
    if target.shape != (B, max_n, c):
        raise RRNetB200Error("target must be [B,max_n,c]")
    loss = torch.empty(1, dtype=torch.float32, device=output.device)
    grad = torch.empty_like(output) if want_grad else None
    check(_lib.lib().rr_regl1_fwd_bwd(_ptr(output), _ptr(mask), _ptr(ind), _ptr(target), B, c, H, W, max_n,
                                      float(grad_scale), _ptr(loss), _ptr(grad), _stream()), "rr_regl1_fwd_bwd")
    return loss, grad


def stage2_loss(bxyxy, seg_offsets, s2_reg, gt_xyxy, scale=4.0, grad_scale=1.0, want_grad=True):
    """Stage-2 regression loss of RRNetOperator.criterion (rrnet_operator.py:64-84) for all images in one launch.
    bxyxy [N,5] image-major, seg_offsets [B+1] int32, s2_reg [N,4], gt_xyxy [B,max_n,>=4] (xyxy, input pixels)
    -> (loss_parts [B], grad_reg [N,4] or None, grad_box [N,4] or None)."""
    bxyxy = _f32(bxyxy, "bxyxy", 2)
    s2_reg = _f32(s2_reg, "s2_reg", 2)
    gt_xyxy = _f32(gt_xyxy, "gt_xyxy", 3)
    seg_offsets = _i32(seg_offsets, "seg_offsets")
    B, max_n, stride = gt_xyxy.shape
    n = bxyxy.shape[0]
    if seg_offsets.numel() != B + 1 or s2_reg.shape != (n, 4) or bxyxy.shape[1] != 5:
        raise RRNetB200Error("stage2_loss: shapes")
    dev = bxyxy.device
    parts = torch.zeros(B, dtype=torch.float32, device=dev)
    g_reg = torch.zeros(n, 4, dtype=torch.float32, device=dev) if want_grad else None
    g_box = torch.zeros(n, 4, dtype=torch.float32, device=dev) if want_grad else None
    check(_lib.lib().rr_stage2_loss(_ptr(bxyxy), _ptr(seg_offsets), _ptr(s2_reg), _ptr(gt_xyxy), B, max_n, stride,
                                    float(scale), float(grad_scale), _ptr(parts), _ptr(g_reg), _ptr(g_box), _stream()),
          "rr_stage2_loss")
    return parts, g_reg, g_box


def ap_match(pred, n_pred, target, n_tgt, thresholds, cls_num=11):
    """Evaluation true-positive matching (utils/metrics/metrics.py:get_tp) for a batch of images.
    pred [B,M,6] x,y,w,h,score,cls; target [B,N,6] x,y,w,h,*,cls (0 = ignore region); n_pred / n_tgt [B] int32
    -> order [B,M] int32, tp [B,M,T], cls [B,M] int32 (-1: not emitted), target_count [B,cls_num-1], in_img [B,cls_num-1]."""
    pred = _f32(pred, "pred", 3)
    target = _f32(target, "target", 3)
    thresholds = _f32(thresholds, "thresholds", 1)
    n_pred = _i32(n_pred, "n_pred")
    n_tgt = _i32(n_tgt, "n_tgt")
    B, M, six = pred.shape
    N, T = target.shape[1], thresholds.numel()
    if six != 6 or target.shape[0] != B or (N and target.shape[2] != 6) or n_pred.numel() != B or n_tgt.numel() != B:
        raise RRNetB200Error("ap_match: pred [B,M,6], target [B,N,6], n_pred / n_tgt [B]")
    dev = pred.device
    order = torch.empty(B, M, dtype=torch.int32, device=dev)
    tp = torch.empty(B, M, T, dtype=torch.float32, device=dev)
    cls = torch.empty(B, M, dtype=torch.int32, device=dev)
    cnt = torch.empty(B, cls_num - 1, dtype=torch.float32, device=dev)
    img = torch.empty(B, cls_num - 1, dtype=torch.float32, device=dev)
    check(_lib.lib().rr_ap_match(_ptr(pred), _ptr(n_pred), _ptr(target), _ptr(n_tgt), _ptr(thresholds), B, M, N, T,
                                 int(cls_num), _ptr(order), _ptr(tp), _ptr(cls), _ptr(cnt), _ptr(img), _stream()),
          "rr_ap_match")
    return order, tp, cls, cnt, img


class KernelTrace:
    """Per-kernel device times of everything this thread launches through the library inside the `with` block
    (rr_kernel_trace_begin / _end): CUDA events between consecutive launches.  So that launch overhead does not show
    up as idle gaps, the stream is first held busy for `hold_ms` (torch.cuda._sleep) while the launches queue up.
    -> .kernels = [(name, ms), ...] after the block (one host sync)."""

    def __init__(self, capacity=64, hold_ms=2.0):
        self.capacity, self.hold_ms = int(capacity), float(hold_ms)
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(self.capacity)]
        self.kernels = []

    def __enter__(self):
        st = torch.cuda.current_stream()
        for e in self.events:
            e.record(st)                                   # torch creates the CUDA event lazily on first record
        if self.hold_ms > 0:
            torch.cuda._sleep(int(self.hold_ms * 1.9e6))   # ~cycles at 1.9 GHz
        self._ev = (ctypes.c_void_p * self.capacity)(*[e.cuda_event for e in self.events])
        self._names = (ctypes.c_char_p * self.capacity)()
        check(_lib.lib().rr_kernel_trace_begin(self._ev, self._names, self.capacity, _stream()), "rr_kernel_trace_begin")
        return self

    def __exit__(self, *exc):
        n = _lib.lib().rr_kernel_trace_end()
        torch.cuda.synchronize()
        self.kernels = [(self._names[i].decode(), self.events[i - 1].elapsed_time(self.events[i])) for i in range(1, n)]
        return False


def set_pdl(enabled):
    """Programmatic dependent launch between the kernels of the eval path (default on)."""
    check(_lib.lib().rr_set_pdl(int(bool(enabled))), "rr_set_pdl")


OPT_PDL, OPT_SELECT_SINGLE_CTA, OPT_COMBINE_IN_TILE_KERNEL = 1, 2, 3


def set_option(option, value):
    """rr_set_option: OPT_PDL, OPT_SELECT_SINGLE_CTA (the decode's selection by one CTA per image instead of a cluster)."""
    check(_lib.lib().rr_set_option(int(option), int(value)), "rr_set_option")


def set_sm_reserve(n_sms):
    """SMs (0..147) the persistent kernels leave free for the short kernels of another batch on another stream.
    Per calling host thread; grid sizes are fixed at launch / graph capture time."""
    check(_lib.lib().rr_set_sm_reserve(int(n_sms)), "rr_set_sm_reserve")


# --------------------------------------------------------------------------------- training side
def render_targets(annos, n_obj, img_h, img_w, scale_factor=4, cls_num=10):
    """to_heatmap + collate padding for a batch: annos [B,max_n,8], n_obj [B] int32 ->
    hm [B,cls,h/sf,w/sf], wh [B,max_n,2], ind [B,max_n,1], offset [B,max_n,2], reg_mask [B,max_n,1]."""
    annos = _f32(annos, "annos", 3)
    n_obj = _i32(n_obj, "n_obj")
    B, max_n, eight = annos.shape
    if eight != 8 or n_obj.numel() != B:
        raise RRNetB200Error("annos must be [B,max_n,8] and n_obj [B]")
    dev = annos.device
    f32 = dict(dtype=torch.float32, device=dev)
    hm = torch.empty(B, cls_num, img_h // scale_factor, img_w // scale_factor, **f32)
    wh = torch.empty(B, max_n, 2, **f32)
    ind = torch.empty(B, max_n, 1, **f32)
    off = torch.empty(B, max_n, 2, **f32)
    msk = torch.empty(B, max_n, 1, **f32)
    check(_lib.lib().rr_render_targets(_ptr(annos), _ptr(n_obj), B, max_n, int(img_h), int(img_w), int(scale_factor),
                                       int(cls_num), _ptr(hm), _ptr(wh), _ptr(ind), _ptr(off), _ptr(msk), _stream()),
          "rr_render_targets")
    return hm, wh, ind, off, msk


def focal_forward(logits, gt):
    """-> stats [4] fp32 on device: loss, pos_sum, neg_sum, num_pos."""
    logits, gt = _f32(logits, "logits"), _f32(gt, "gt")
    if logits.numel() != gt.numel():
        raise RRNetB200Error("logits and gt must have the same number of elements")
    L = _lib.lib()
    stats = torch.empty(4, dtype=torch.float32, device=logits.device)
    ws = _ws(L.rr_focal_workspace_bytes(logits.numel()), logits.device)
    check(L.rr_focal_forward(_ptr(logits), _ptr(gt), logits.numel(), _ptr(stats), _ptr(ws), ws.numel(), _stream()),
          "rr_focal_forward")
    return stats


def focal_backward(logits, gt, stats, upstream=1.0):
    logits, gt = _f32(logits, "logits"), _f32(gt, "gt")
    grad = torch.empty_like(logits)
    check(_lib.lib().rr_focal_backward(_ptr(logits), _ptr(gt), logits.numel(), _ptr(_f32(stats, "stats")),
                                       float(upstream), _ptr(grad), _stream()), "rr_focal_backward")
    return grad


def focal_fwd_bwd(logits, gt, upstream=1.0):
    """Loss and d loss/d logits in one cooperative launch -> (stats [4], grad)."""
    logits, gt = _f32(logits, "logits"), _f32(gt, "gt")
    L = _lib.lib()
    stats = torch.empty(4, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits)
    ws = _ws(L.rr_focal_workspace_bytes(logits.numel()), logits.device)
    check(L.rr_focal_fwd_bwd(_ptr(logits), _ptr(gt), logits.numel(), float(upstream), _ptr(stats), _ptr(grad),
                             _ptr(ws), ws.numel(), _stream()), "rr_focal_fwd_bwd")
    return stats, grad


def focal_render_forward(logits, annos, n_obj, img_h, img_w, scale_factor=4, want_gt=False):
    """Heat-map focal loss straight from the padded annotations (the target map is rendered on the fly and
    never stored).  logits [B,cls,h,w], annos [B,max_n,8], n_obj [B] int32 -> stats [4] (+ gt map if want_gt)."""
    logits, annos, n_obj = _f32(logits, "logits", 4), _f32(annos, "annos", 3), _i32(n_obj, "n_obj")
    B, C, h, w = logits.shape
    if h != img_h // scale_factor or w != img_w // scale_factor or annos.shape[0] != B or annos.shape[2] != 8:
        raise RRNetB200Error("logits must be [B,cls,img_h/sf,img_w/sf] and annos [B,max_n,8]")
    L = _lib.lib()
    stats = torch.empty(4, dtype=torch.float32, device=logits.device)
    gt = torch.empty_like(logits) if want_gt else None
    ws = _ws(L.rr_focal_render_workspace_bytes(B, C, int(img_h), int(img_w), int(scale_factor)), logits.device)
    check(L.rr_focal_render_forward(_ptr(logits), _ptr(annos), _ptr(n_obj), B, annos.shape[1], int(img_h), int(img_w),
                                    int(scale_factor), C, _ptr(stats), _ptr(gt), _ptr(ws), ws.numel(), _stream()),
          "rr_focal_render_forward")
    return (stats, gt) if want_gt else stats


def focal_render_fwd_bwd(logits, annos, n_obj, img_h, img_w, upstream=1.0, scale_factor=4):
    """Loss AND gradient of the heat-map focal term from the padded annotations in one pass over the logits (the target
    is rendered tile by tile in shared memory).  -> (stats [4] = loss, pos_sum, neg_sum, num_pos; grad like logits)."""
    logits, annos, n_obj = _f32(logits, "logits", 4), _f32(annos, "annos", 3), _i32(n_obj, "n_obj")
    B, C, h, w = logits.shape
    if h != img_h // scale_factor or w != img_w // scale_factor or annos.shape[0] != B or annos.shape[2] != 8:
        raise RRNetB200Error("logits must be [B,cls,img_h/sf,img_w/sf] and annos [B,max_n,8]")
    L = _lib.lib()
    stats = torch.empty(4, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits)
    ws = _ws(L.rr_focal_render_workspace_bytes(B, C, int(img_h), int(img_w), int(scale_factor)), logits.device)
    check(L.rr_focal_render_fwd_bwd(_ptr(logits), _ptr(annos), _ptr(n_obj), B, annos.shape[1], int(img_h), int(img_w),
                                    int(scale_factor), C, float(upstream), _ptr(stats), _ptr(grad), _ptr(ws), ws.numel(),
                                    _stream()), "rr_focal_render_fwd_bwd")
    return stats, grad


def focal_render_backward(logits, annos, n_obj, img_h, img_w, stats, upstream=1.0, scale_factor=4):
    logits, annos, n_obj = _f32(logits, "logits", 4), _f32(annos, "annos", 3), _i32(n_obj, "n_obj")
    B, C, h, w = logits.shape
    grad = torch.empty_like(logits)
    check(_lib.lib().rr_focal_render_backward(_ptr(logits), _ptr(annos), _ptr(n_obj), B, annos.shape[1], int(img_h),
                                              int(img_w), int(scale_factor), C, _ptr(_f32(stats, "stats")),
                                              float(upstream), _ptr(grad), _stream()), "rr_focal_render_backward")
    return grad
