"""ext/nms/nms/py_cpu_nms.py:4-32 -- "+1" areas, keeps IoU <= thresh (suppress iff IoU > thresh)."""
import numpy as np
import torch

from rrnet_b200 import ops


def py_cpu_nms(dets, thresh):
    dets = np.asarray(dets)
    if dets.shape[0] == 0:
        return []
    order = dets[:, 4].argsort()[::-1]
    d = torch.from_numpy(np.ascontiguousarray(dets[order, :5], dtype=np.float32)).cuda()
    keep = ops.nms(d[:, :4].contiguous(), d[:, 4].contiguous(), float(thresh), pixel_offset=1, ge_cmp=False)
    return list(order[keep.cpu().numpy()])
