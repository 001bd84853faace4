"""The reference's CPU call sequence for the post-backbone path, as a timed baseline.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is what `bench.py`'s `cpu_baseline`
leg and `bench.py --impl reference` execute on the GPU box's host cores: the same third-party
CPU kernels the reference calls -- torch.sigmoid / torch.topk / gather (models/rrnet.py:93-138),
torchvision.ops.nms per class (:56-72), torch.relu + torchvision.ops.roi_align (:51), the
Bottleneck head through torch's CPU convolutions (detectors/fasterrcnn_detector.py:13-18,
backbones/resnet.py:33-53), the elementwise box decode (operators/rrnet_operator.py:188-209) --
and, for the final stage, the reference's own Cython soft-NMS compiled into oracle/_ref
(ext/nms/nms/cpu_nms.pyx:17-120).  /root/reference itself does not exist on the GPU box, so the
call sequence is restated here; tests/test_ref_port.py checks it against the golden fixtures that
the unmodified reference produced.
"""
import numpy as np
import torch
import torch.nn.functional as F
import torchvision


def topk_decode(hm, wh, off, K):
    """[B,C,H,W] logits -> [B,K,6] rows x1,y1,x2,y2,score,cls (stride-4 units).
    Two-stage top-K exactly as the reference does it: per class over H*W, then over C*K."""
    B, C, H, W = hm.shape
    heat = torch.sigmoid(hm)
    per_cls_score, per_cls_ind = torch.topk(heat.reshape(B, C, H * W), K)           # rrnet.py:96
    per_cls_ind = per_cls_ind % (H * W)
    per_cls_y = (per_cls_ind / W).int().float()                                      # :99
    per_cls_x = (per_cls_ind % W).int().float()                                      # :100
    score, pick = torch.topk(per_cls_score.reshape(B, C * K), K)                     # :102
    cls = (pick / K).int()                                                           # :103
    ind = per_cls_ind.reshape(B, C * K).gather(1, pick)
    ys = per_cls_y.reshape(B, C * K).gather(1, pick)
    xs = per_cls_x.reshape(B, C * K).gather(1, pick)

    def at(m):                                                                       # :111-115
        flat = m.permute(0, 2, 3, 1).contiguous().reshape(B, H * W, m.shape[1])
        return flat.gather(1, ind.unsqueeze(2).expand(B, K, m.shape[1]))

    o = at(off)
    size = at(wh).clamp(min=0)                                                       # :128
    cx = xs.unsqueeze(2) + o[..., 0:1]
    cy = ys.unsqueeze(2) + o[..., 1:2]
    x1 = cx - size[..., 0:1] / 2
    y1 = cy - size[..., 1:2] / 2
    rows = torch.cat([x1, y1, size[..., 0:1] + x1, size[..., 1:2] + y1,
                      score.unsqueeze(2), cls.float().unsqueeze(2)], dim=2)          # :137
    return rows, ind


def per_class_nms(rows, thr=0.7):
    """One image's [K,6] rows -> kept rows, class ascending, score descending (rrnet.py:56-72)."""
    out = []
    for c in rows[:, 5].unique():
        sel = rows[rows[:, 5] == c]
        keep = torchvision.ops.nms(sel[:, :4], sel[:, 4], thr)
        out.append(sel[keep])
    return torch.cat(out) if out else rows[:0]


def make_head(hp):
    """nn.Module-free Bottleneck(256,64)+regressor in eval mode from synth.head_params() tensors."""
    def bn(x, p):
        return F.batch_norm(x, p[2], p[3], p[0], p[1], training=False, eps=1e-5)

    def head(x):
        t = F.relu(bn(F.conv2d(x, hp["w1"].view(64, 256, 1, 1)), hp["bn1"]))
        t = F.relu(bn(F.conv2d(t, hp["w2"], padding=1), hp["bn2"]))
        t = bn(F.conv2d(t, hp["w3"].view(256, 64, 1, 1)), hp["bn3"])
        t = F.relu(t + x)
        t = F.adaptive_avg_pool2d(t, 1)
        return F.conv2d(t, hp["wr"].view(4, 256, 1, 1), hp["br"]).flatten(1)
    return head


def box_decode(bxyxy, reg, scores, clses, scale=4.0):
    """rrnet_operator.py:188-209 for all rows -> (s1 [n,6], s2 [n,6])."""
    xy = bxyxy[:, 1:3] * scale
    wh = bxyxy[:, 3:5] * scale - xy
    s1 = torch.cat([xy, wh, scores[:, None], torch.zeros_like(scores)[:, None]], 1)
    wh1 = wh + 1
    ctr = reg[:, 0:2] * wh1 + xy + wh1 / 2
    out_wh = reg[:, 2:4].exp() * wh1
    s2 = torch.cat([ctr - out_wh / 2.0, out_wh, scores[:, None], clses[:, None] + 1], 1)
    return s1, s2


@torch.no_grad()
def post_backbone(hm, wh, off, feat, hp, K, nms_thr=0.7, scale=4.0, stage_times=None):
    """RRNet.forward after forward_stage1 + generate_bbox, on CPU tensors.  Returns dict like
    rrnet_b200.ops.EvalPath.results().  stage_times (dict) accumulates per-stage seconds."""
    import time
    t0 = time.perf_counter()
    rows, _ = topk_decode(hm, wh, off, K)
    t1 = time.perf_counter()
    kept = [per_class_nms(rows[b], nms_thr) for b in range(rows.shape[0])]
    counts = [k.shape[0] for k in kept]
    bxyxy = torch.cat([torch.cat([torch.full((k.shape[0], 1), float(b), device=k.device), k[:, :4]], 1) for b, k in enumerate(kept)])
    scores = torch.cat([k[:, 4] for k in kept])
    clses = torch.cat([k[:, 5] for k in kept])
    t2 = time.perf_counter()
    roi = torchvision.ops.roi_align(torch.relu(feat), bxyxy, (3, 3))                 # rrnet.py:51
    t3 = time.perf_counter()
    reg = make_head(hp)(roi)
    t4 = time.perf_counter()
    s1, s2 = box_decode(bxyxy, reg, scores, clses, scale)
    t5 = time.perf_counter()
    if stage_times is not None:
        for k, v in (("decode", t1 - t0), ("nms", t2 - t1), ("relu_roi_align", t3 - t2), ("head", t4 - t3),
                     ("bbox", t5 - t4)):
            stage_times[k] = stage_times.get(k, 0.0) + v
    return {"n": bxyxy.shape[0], "counts": counts, "bxyxy": bxyxy, "scores": scores, "clses": clses,
            "reg": reg, "s1": s1, "s2": s2, "dets": rows}


def final_soft_nms(s2, cpu_nms_mod):
    """RRNetOperator._ext_nms (rrnet_operator.py:211-232) with the reference's compiled cpu_soft_nms."""
    if s2.shape[0] == 0:
        return s2
    out = []
    for c in s2[:, 5].unique():
        rows = s2[s2[:, 5] == c].numpy().copy()
        rows[:, 2] += rows[:, 0]
        rows[:, 3] += rows[:, 1]
        keep = cpu_nms_mod.cpu_soft_nms(np.ascontiguousarray(rows, dtype=np.float32), np.float32(0.5),
                                        np.float32(0.7), np.float32(0.1), np.uint8(2))
        out.append(rows[keep])
    out = np.concatenate(out, 0)
    out[:, 2:4] -= out[:, 0:2]
    return torch.from_numpy(out)


def focal_loss(logits, gt):
    """criterion's heat-map term (rrnet_operator.py:55-57 + modules/loss/functional.py:25-51)."""
    p = torch.clamp(torch.sigmoid(logits), min=1e-4, max=1 - 1e-4)
    pos = gt.eq(1).float()
    neg = gt.lt(1).float()
    pos_loss = (torch.log(p) * torch.pow(1 - p, 2) * pos).sum()
    neg_loss = (torch.log(1 - p) * torch.pow(p, 2) * torch.pow(1 - gt, 4) * neg).sum()
    n = pos.sum()
    return -neg_loss if n == 0 else -(pos_loss + neg_loss) / n
