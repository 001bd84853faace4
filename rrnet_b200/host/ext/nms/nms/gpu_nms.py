"""ext/nms/nms/gpu_nms.pyx:16-31 -- hard NMS on the GPU, "+1" areas, suppress iff IoU > thresh."""
import numpy as np

from rrnet_b200 import ops


def gpu_nms(dets, thresh, device_id=0):
    """dets: numpy [n, >=5] (x1,y1,x2,y2,score) on the host -> list of kept row indices, score
    descending.  As the reference: rows are ordered on the host with `scores.argsort()[::-1]`
    (gpu_nms.pyx:25), the sorted rows go through the `_nms` ABI (rr_nms_legacy_host replaces
    nms_kernel.cu:91-144; the mask is reduced on the device instead of being copied back)."""
    dets = np.asarray(dets)
    boxes_num = dets.shape[0]
    if boxes_num == 0:
        return []
    scores = dets[:, 4]
    order = scores.argsort()[::-1]
    sorted_dets = np.ascontiguousarray(dets[order, :5], dtype=np.float32)
    keep = ops.nms_legacy_host(sorted_dets, np.float32(thresh), device_id)
    return list(order[keep])
