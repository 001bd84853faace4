"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the per-image true-positive matching of the reference's
evaluation, utils/metrics/metrics.py:10-47 (`bbox_iou`) and :51-136 (`get_tp`).  Imported by tests/ only; the product
(`rrnet_b200/`) never imports anything under oracle/.

Pinned against the reference itself: tests/golden/ap_match.npz holds the outputs of the UNMODIFIED `get_tp` on seeded
images (tests/golden/make_golden.py:gold_ap_match), tests/test_oracle_golden.py checks this file against them.

All arithmetic in float32, element for element like the torch CPU ops the reference calls."""
import numpy as np

F = np.float32


def bbox_iou_xywh(a, b):
    """metrics.py:10-47 with x1y1x2y2=False, overlap=True: -> (IoU [m,n], inter / area_a [m,n])."""
    a = a.astype(F).copy()
    b = b.astype(F).copy()
    a[:, 2] += a[:, 0]; a[:, 3] += a[:, 1]                       # :22-26
    b[:, 2] += b[:, 0]; b[:, 3] += b[:, 1]
    a_area = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])            # :28-29
    b_area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iw = np.minimum(a[:, 2][:, None], b[:, 2]) - np.maximum(a[:, 0][:, None], b[:, 0])   # :31-32
    ih = np.minimum(a[:, 3][:, None], b[:, 3]) - np.maximum(a[:, 1][:, None], b[:, 1])
    iw = np.maximum(iw, F(0)); ih = np.maximum(ih, F(0))          # :34-35
    ua = a_area[:, None] + b_area - iw * ih                       # :37
    ua = np.maximum(ua, F(1e-8))                                  # :39
    inter = iw * ih
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / ua, inter / a_area[:, None]                # :41-45


def get_tp_image(pred, target, thresholds, cls_num=11):
    """metrics.py:51-136 for ONE image, without the cross-image concatenation.
    pred [m,6] = x,y,w,h,score,cls ; target [n,6] = x,y,w,h,*,cls (cls 0 = ignore region)
    -> dict: order [m'] (indices into pred, score-descending, detections inside ignore regions removed),
             tp [m',T] (0/1 per IoU threshold), conf [m'], cls [m'], emit [m'] (the reference appends a detection to
             its class lists only if the image has ground truth of that class, :113-114), target_count [cls_num-1],
             in_img [cls_num-1]."""
    pred = np.asarray(pred, F)
    target = np.asarray(target, F)
    thr = np.asarray(thresholds, F)
    T = thr.shape[0]
    order = np.argsort(-pred[:, 4], kind="stable")                # :70-71 (ties: the goldens have none)
    pred = pred[order]
    ignore = target[:, 5] == 0                                    # :74-79
    if ignore.sum() != 0:
        _, gt_ov = bbox_iou_xywh(target[:, :4], target[:, :4])
        keep = (gt_ov[:, ignore].max(axis=1) < F(0.5)) | ignore
        target = target[keep]
    ignore = target[:, 5] == 0                                    # :82-88
    iou, ov = bbox_iou_xywh(pred[:, :4], target[:, :4])
    if ignore.sum() != 0:
        keep = ov[:, ignore].max(axis=1) < F(0.5)
        pred, iou, order = pred[keep], iou[keep], order[keep]
    pcls = pred[:, 5].astype(np.int64)
    tcls = target[:, 5].astype(np.int64)
    same = pcls[:, None] == tcls[None, :]                         # :93-95
    flag = (iou[:, :, None] - thr[None, None, :]) >= 0            # :97
    tp_iou = iou[:, :, None] * (same[:, :, None] & flag).astype(F)   # :99-101
    tp = np.zeros((pred.shape[0], T), F)
    emit = np.zeros(pred.shape[0], bool)
    target_count = np.zeros(cls_num - 1, F)
    in_img = np.zeros(cls_num - 1, F)
    for c in range(1, cls_num):                                   # :105-134
        d_idx = np.nonzero(pcls == c)[0]
        g_idx = np.nonzero(tcls == c)[0]
        target_count[c - 1] += g_idx.shape[0]
        in_img[c - 1] += 1 if g_idx.shape[0] != 0 else 0
        if d_idx.shape[0] == 0 or g_idx.shape[0] == 0:
            continue
        emit[d_idx] = True
        m = tp_iou[np.ix_(d_idx, g_idx)].copy()                   # [d, g, T]
        for i in range(d_idx.shape[0]):
            best = m[i].max(axis=0)                               # per threshold: max over the ground truth
            arg = m[i].argmax(axis=0)                             # first maximum, like torch.max
            for t in np.nonzero(best)[0]:
                m[:, arg[t], t] = 0                               # that box is used up at that threshold
                tp[d_idx[i], t] = 1
    return {"order": order, "tp": tp, "conf": pred[:, 4].copy(), "cls": pcls.astype(np.int32), "emit": emit,
            "target_count": target_count, "in_img": in_img}
