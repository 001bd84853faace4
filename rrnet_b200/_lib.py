"""ctypes binding of librrnet_b200.so (the C ABI declared in include/rrnet_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module
raises.  `python -m rrnet_b200.build` (or `__graft_entry__.build()`) produces the library.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librrnet_b200.so")

c_int, c_float, c_double = ctypes.c_int, ctypes.c_float, ctypes.c_double
c_size_t, c_void_p, c_int64 = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int64
P = c_void_p  # every pointer crosses the boundary as a raw address

# name -> (restype, argtypes); mirrors include/rrnet_b200.h one to one
SIGNATURES = {
    "rr_version": (c_int, []),
    "rr_error_string": (ctypes.c_char_p, [c_int]),
    "rr_launch_count": (ctypes.c_uint64, []),
    "rr_set_sm_reserve": (c_int, [c_int]),
    "rr_set_pdl": (c_int, [c_int]),
    "rr_set_option": (c_int, [c_int, c_int]),
    "rr_kernel_trace_begin": (c_int, [P, P, c_int, P]),
    "rr_kernel_trace_end": (c_int, []),
    "rr_decode_workspace_bytes": (c_size_t, [c_int] * 5),
    "rr_decode_topk": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "rr_hm_tail_collect": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_size_t, P]),
    "rr_stage1_nms_workspace_bytes": (c_size_t, [c_int] * 3),
    "rr_stage1_nms": (c_int, [P, c_int, c_int, c_int, c_double, P, P, P, P, P, c_size_t, P]),
    "rr_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rr_nms_batched": (c_int, [P, P, P, c_int, c_int, c_double, c_int, c_int, P, P, P, c_size_t, P]),
    "rr_nms_legacy_host": (c_int, [P, P, P, c_int, c_int, c_float, c_int]),
    "rr_soft_nms_workspace_bytes": (c_size_t, [c_int]),
    "rr_soft_nms_batched": (c_int, [P, P, c_int, c_int, c_float, c_float, c_float, c_int, P, P, P, c_size_t, P]),
    "rr_roi_align_workspace_bytes": (c_size_t, [c_int] * 5),
    "rr_roi_align": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_size_t, P]),
    "rr_roi_align_backward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "rr_head_folded_floats": (c_size_t, []),
    "rr_head_fold": (c_int, [P] * 9 + [P]),
    "rr_head_forward": (c_int, [P, P, c_int, P, c_int, P, P]),
    "rr_generate_bbox": (c_int, [P, P, P, P, P, c_int, c_float, P, P, P]),
    "rr_eval_workspace_bytes": (c_size_t, [c_int] * 6),
    "rr_eval_forward": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double,
                                c_int, P, c_float, P, P, P, P, P, P, P, P, P, P, P, c_size_t, P, P]),
    "rr_render_targets": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P]),
    "rr_focal_workspace_bytes": (c_size_t, [c_int64]),
    "rr_focal_forward": (c_int, [P, P, c_int64, P, P, c_size_t, P]),
    "rr_focal_backward": (c_int, [P, P, c_int64, P, c_float, P, P]),
    "rr_focal_fwd_bwd": (c_int, [P, P, c_int64, c_float, P, P, P, c_size_t, P]),
    "rr_focal_render_workspace_bytes": (c_size_t, [c_int] * 5),
    "rr_focal_render_forward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "rr_focal_render_fwd_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P, c_size_t, P]),
    "rr_ap_match": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P]),
    "rr_stage2_loss": (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, c_float, P, P, P, P]),
    "rr_regl1_fwd_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P]),
    "rr_focal_render_backward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_float, P, P]),
}

_lib = None


class RRNetB200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RRNetB200Error(
                "librrnet_b200.so not found at %s -- build it with `python -m rrnet_b200.build`; "
                "there is no CPU fallback" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().rr_error_string(rc).decode()
        raise RRNetB200Error("%s failed: %s (code %d)" % (what, msg, rc))


def launch_count():
    return int(lib().rr_launch_count())
