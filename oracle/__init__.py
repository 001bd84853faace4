"""CPU oracle for the RRNet post-backbone hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  The product (``rrnet_b200``) never imports it and has no CPU
fallback.

``oracle.rr_oracle.c`` is the C restatement (numpy in / numpy out wrappers below);
``oracle.ref_port`` is the reference's torch-CPU call sequence used as the CPU
baseline; ``oracle.build_ref`` compiles the reference's own Cython NMS into
``oracle/_ref``.  Parity pinning is described in ``rr_oracle.c``'s header and DESIGN.md.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile oracle/_build/liboracle.so with the Makefile next to this file."""
    src = os.path.join(_HERE, "rr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_version.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def decode(hm, wh, off, K, pool=0):
    """-> dets [B,K,6] f32, inds [B,K] i64 (y*W+x), flat [B,K] i64 (cls*H*W + ind)."""
    hm, wh, off = _f32(hm), _f32(wh), _f32(off)
    B, C, H, W = hm.shape
    dets = np.empty((B, K, 6), np.float32)
    inds = np.empty((B, K), np.int64)
    flat = np.empty((B, K), np.int64)
    rc = lib().orc_decode(_p(hm, _f32p), _p(wh, _f32p), _p(off, _f32p), B, C, H, W, int(K), int(pool),
                          _p(dets, _f32p), _p(inds, _i64p), _p(flat, _i64p))
    if rc != 0:
        raise ValueError("orc_decode rc=%d (K must satisfy 0 < K <= H*W)" % rc)
    return dets, inds, flat


def nms(boxes, scores, thr, pixel_offset=0, ge_cmp=False):
    """Greedy hard NMS; returns original indices in acceptance order (int32)."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), np.int32)
    lib().orc_nms.restype = ctypes.c_int
    nk = lib().orc_nms(_p(boxes, _f32p), _p(scores, _f32p), n, ctypes.c_double(thr),
                       int(pixel_offset), int(bool(ge_cmp)), _p(keep, _i32p))
    return keep[:nk].copy()


def nms_sorted(boxes, thr, pixel_offset=1, ge_cmp=False):
    """Greedy NMS over rows that are already score-sorted (legacy _nms ABI layout)."""
    boxes = _f32(boxes)
    n, dim = boxes.shape
    keep = np.empty(max(n, 1), np.int32)
    nk = lib().orc_nms_sorted(_p(boxes, _f32p), n, dim, ctypes.c_double(thr), int(pixel_offset),
                              int(bool(ge_cmp)), _p(keep, _i32p))
    return keep[:nk].copy()


def stage1_nms(dets, num_classes=10, thr=0.7):
    """RRNet.nms default branch on one image's [K,6] rows -> (kept rows [n,6], source rows [n])."""
    dets = _f32(dets)
    K = dets.shape[0]
    out = np.empty((max(K, 1), 6), np.float32)
    src = np.empty(max(K, 1), np.int32)
    n = lib().orc_stage1_nms(_p(dets, _f32p), K, int(num_classes), ctypes.c_double(thr),
                             _p(out, _f32p), _p(src, _i32p))
    return out[:n].copy(), src[:n].copy()


def soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=1):
    """cpu_soft_nms on a COPY of boxes [n,5]; returns the kept, re-scored rows [N,5]."""
    b = _f32(boxes)[:, :5].copy()
    N = lib().orc_soft_nms(_p(b, _f32p), b.shape[0], ctypes.c_float(sigma), ctypes.c_float(Nt),
                           ctypes.c_float(threshold), int(method))
    return b[:N].copy()


def roi_align(feat, rois, pooled=(3, 3), relu=True):
    feat, rois = _f32(feat), _f32(rois).reshape(-1, 5)
    B, C, H, W = feat.shape
    N = rois.shape[0]
    out = np.empty((N, C, pooled[0], pooled[1]), np.float32)
    rc = lib().orc_roi_align(_p(feat, _f32p), _p(rois, _f32p), N, B, C, H, W, pooled[0], pooled[1],
                             int(bool(relu)), _p(out, _f32p))
    if rc != 0:
        raise ValueError("orc_roi_align: roi batch index out of range")
    return out


HEAD_KEYS = ("w1", "bn1", "w2", "bn2", "w3", "bn3", "wr", "br")


def head(x, params):
    """params: dict with HEAD_KEYS; bnX = [4,ch] rows gamma,beta,running_mean,running_var."""
    x = _f32(x)
    N = x.shape[0]
    assert x.shape[1:] == (256, 3, 3)
    p = {k: _f32(params[k]) for k in HEAD_KEYS}
    out = np.empty((N, 4), np.float32)
    lib().orc_head(_p(x, _f32p), N, *[_p(p[k], _f32p) for k in HEAD_KEYS], _p(out, _f32p))
    return out


def generate_bbox(bxyxy, reg, scores, clses, batch_idx=0, scale=4.0):
    bxyxy, reg, scores, clses = _f32(bxyxy).reshape(-1, 5), _f32(reg).reshape(-1, 4), _f32(scores), _f32(clses)
    N = bxyxy.shape[0]
    s1 = np.empty((max(N, 1), 6), np.float32)
    s2 = np.empty((max(N, 1), 6), np.float32)
    m = lib().orc_generate_bbox(_p(bxyxy, _f32p), _p(reg, _f32p), _p(scores, _f32p), _p(clses, _f32p),
                                N, int(batch_idx), ctypes.c_float(scale), _p(s1, _f32p), _p(s2, _f32p))
    return s1[:m].copy(), s2[:m].copy()


def render(annos, img_h, img_w, scale_factor=4, cls_num=10, hm=None):
    """to_heatmap for one image: annos [n,8] -> dict(hm, wh, ind, offset, reg_mask, radius)."""
    annos = _f32(annos).reshape(-1, 8)
    n = annos.shape[0]
    Hh, Wh = img_h // scale_factor, img_w // scale_factor
    if hm is None:
        hm = np.zeros((cls_num, Hh, Wh), np.float32)
    wh = np.zeros((n, 2), np.float32)
    ind = np.zeros((n, 1), np.float32)
    off = np.zeros((n, 2), np.float32)
    msk = np.zeros((n, 1), np.float32)
    rad = np.zeros((n,), np.float32)
    rc = lib().orc_render(_p(annos, _f32p), n, int(img_h), int(img_w), int(scale_factor), int(cls_num),
                          _p(hm, _f32p), _p(wh, _f32p), _p(ind, _f32p), _p(off, _f32p), _p(msk, _f32p),
                          _p(rad, _f32p))
    if rc != 0:
        raise ValueError("orc_render: class index out of range")
    return dict(hm=hm, wh=wh, ind=ind, offset=off, reg_mask=msk, radius=rad)


def focal(logits, gt, want_grad=False):
    """-> (loss float64, sums [pos, neg, num_pos] float64, grad f32 or None)."""
    logits, gt = _f32(logits), _f32(gt)
    n = logits.size
    sums = np.zeros(3, np.float64)
    loss = ctypes.c_double(0)
    grad = np.empty(logits.shape, np.float32) if want_grad else None
    lib().orc_focal(_p(logits, _f32p), _p(gt, _f32p), ctypes.c_int64(n), _p(sums, _f64p),
                    ctypes.byref(loss), _p(grad, _f32p))
    return loss.value, sums, grad
