"""ext/nms/nms/cpu_nms.pyx -- same names and semantics, computed on the GPU (no CPU path here).

cpu_nms      (:122-173)  "+1" areas, suppress iff IoU >= thresh (double threshold), order = argsort()[::-1]
cpu_soft_nms (:17-120)   linear / gaussian / hard soft-NMS, IN PLACE on `boxes`, returns list(range(N))
"""
import numpy as np
import torch

from rrnet_b200 import ops


def cpu_nms(dets, thresh):
    dets = np.asarray(dets)
    n = dets.shape[0]
    if n == 0:
        return []
    order = dets[:, 4].argsort()[::-1]
    d = torch.from_numpy(np.ascontiguousarray(dets[order, :5], dtype=np.float32)).cuda()
    # rows are already score-descending; the device sort is stable, so ties keep this order
    keep = ops.nms(d[:, :4].contiguous(), d[:, 4].contiguous(), float(thresh), pixel_offset=1, ge_cmp=True)
    return list(order[keep.cpu().numpy()])


def cpu_soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """boxes: C-contiguous float32 numpy [N, >=5]; columns 0..4 are reordered / decayed IN PLACE exactly
    like the reference (selection by current maximum, swap with the last row on removal; columns >= 5
    are not moved, cpu_nms.pyx:54-66).  Returns [0 .. N'-1]."""
    if not (isinstance(boxes, np.ndarray) and boxes.dtype == np.float32 and boxes.flags["C_CONTIGUOUS"]):
        raise TypeError("cpu_soft_nms expects a C-contiguous float32 ndarray (Cython buffer signature)")
    n = boxes.shape[0]
    if n == 0:
        return []
    d = torch.from_numpy(np.ascontiguousarray(boxes[:, :5])).cuda()
    seg = torch.tensor([0, n], dtype=torch.int32, device=d.device)
    rows, _, cnt = ops.soft_nms_batched(d, seg, float(sigma), float(Nt), float(threshold), int(method))
    boxes[:, :5] = rows.cpu().numpy()
    return list(range(int(cnt.item())))
