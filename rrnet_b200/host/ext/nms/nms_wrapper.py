"""ext/nms/nms_wrapper.py -- soft_nms (:13-19) and nms (:23-33), numpy in / numpy rows out."""
import numpy as np

from .nms.cpu_nms import cpu_nms, cpu_soft_nms
from .nms.gpu_nms import gpu_nms


def soft_nms(dets, sigma=0.5, Nt=0.3, threshold=0.001, method=1):
    """Returns dets[keep].  As in the reference, `np.ascontiguousarray(dets, dtype=np.float32)` is the
    array that gets reordered in place: when dets already is C-contiguous float32 (every call site of
    the reference: .cpu().numpy() rows) that is dets itself, so the returned rows carry the decayed
    scores; any other input is left untouched and its first N' rows come back (reference quirk)."""
    keep = cpu_soft_nms(np.ascontiguousarray(dets, dtype=np.float32),
                        np.float32(sigma), np.float32(Nt),
                        np.float32(threshold),
                        np.uint8(method))
    results = dets[keep]
    return results


def nms(dets, thresh, gpu_id=0):
    """Returns the kept ROWS (not indices); [] for empty input.  gpu_id=None selects the reference's
    CPU semantics (IoU >= thresh), still evaluated on the current CUDA device."""
    if dets.shape[0] == 0:
        return []
    else:
        if gpu_id is not None:
            keep = gpu_nms(dets[:, :5], thresh, device_id=gpu_id)
            results = dets[keep]
            return results
        keep = cpu_nms(dets[:, :5], thresh)
        results = dets[keep]
        return results
