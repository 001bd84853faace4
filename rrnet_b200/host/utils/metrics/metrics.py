"""utils/metrics/metrics.py:51-136 -- get_tp on the GPU.

Same signature and accumulation protocol as the reference: the per-class lists `cls_tp_flags` / `cls_tp_confs` grow by
the image's emitted detections, the two count tensors are incremented.  `get_tp_batch` does a whole batch of images
with one launch (rr_ap_match) and one device->host copy.  The AP / recall integration (`calculate_ap_rc`) is a few
cumulative sums over those lists and stays the reference's."""
import torch

from .... import ops

_THRESHOLDS = torch.arange(0.5, 1.0, 0.05)


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def get_tp_batch(preds, targets, thresholds=_THRESHOLDS, cls_num=11):
    """preds / targets: lists of [m_i,6] / [n_i,6] tensors (any device) ->
    per image: (tp [k_i,T], conf [k_i], cls [k_i]) for the emitted detections in score order, plus
    target_count [B,cls_num-1] and in_img [B,cls_num-1] (CPU tensors)."""
    B = len(preds)
    dev = _device()
    M = max(1, max(int(p.shape[0]) for p in preds))
    N = max(1, max(int(t.shape[0]) for t in targets))
    pred = torch.zeros(B, M, 6)
    tgt = torch.zeros(B, N, 6)
    for b in range(B):
        pred[b, : preds[b].shape[0]] = preds[b].detach().float().cpu()[:, :6]
        tgt[b, : targets[b].shape[0]] = targets[b].detach().float().cpu()[:, :6]
    n_pred = torch.tensor([int(p.shape[0]) for p in preds], dtype=torch.int32)
    n_tgt = torch.tensor([int(t.shape[0]) for t in targets], dtype=torch.int32)
    order, tp, cls, cnt, img = ops.ap_match(pred.to(dev), n_pred.to(dev), tgt.to(dev), n_tgt.to(dev),
                                           thresholds.float().to(dev), cls_num)
    order, tp, cls, cnt, img = order.cpu(), tp.cpu(), cls.cpu(), cnt.cpu(), img.cpu()
    out = []
    for b in range(B):
        m = int(n_pred[b])
        keep = cls[b, :m] >= 0
        conf = pred[b, order[b, :m].long(), 4]
        out.append((tp[b, :m][keep], conf[keep], cls[b, :m][keep]))
    return out, cnt, img


def get_tp(pred, target, cls_tp_flags, cls_tp_confs, cls_target_count, cls_in_img_count,
           thresholds=_THRESHOLDS, cls_num=11):
    """metrics.py:51-136, one image."""
    per_image, cnt, img = get_tp_batch([pred], [target], thresholds, cls_num)
    tp, conf, cls = per_image[0]
    for c in range(1, cls_num):
        sel = cls == c
        if int(sel.sum()) == 0:
            continue
        cls_tp_flags[c - 1] = torch.cat((cls_tp_flags[c - 1], tp[sel]))
        cls_tp_confs[c - 1] = torch.cat((cls_tp_confs[c - 1], conf[sel]))
    cls_target_count += cnt[0]
    cls_in_img_count += img[0]
    return cls_tp_flags, cls_tp_confs, cls_target_count, cls_in_img_count
