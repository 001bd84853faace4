"""The C-ABI library loads on a CPU-only box and exports every symbol include/rrnet_b200.h declares.
No compute is issued here (no GPU)."""
import ctypes
import os
import re

import pytest

from tests.conftest import REPO


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "rrnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from rrnet_b200 import build as rr_build
    path = rr_build.build()
    handle = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 24
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, "declared in include/rrnet_b200.h but not exported: %s" % missing


def test_binding_table_matches_header():
    from rrnet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    L = _lib.lib()
    assert L.rr_version() >= 100
    assert b"workspace" in L.rr_error_string(-2)
    # size queries are pure host code
    assert L.rr_head_folded_floats() == (256 * 64 + 64 + 9 * 64 * 64 + 64 + 64 * 256 + 256 + 4 * 256 + 4) + 17 * 2 * 64 * 32      # tensor-core image: 17 steps of (hi | lo) 64 x 128-byte fp16 tiles
    assert L.rr_decode_workspace_bytes(8, 10, 272, 480, 1500) >= 8 * 16384 * 8
    assert L.rr_eval_workspace_bytes(8, 10, 272, 480, 1500, 256) > 8 * 1500 * 256 * 9 * 4


def test_argument_errors_are_reported_without_a_gpu():
    from rrnet_b200 import _lib
    L = _lib.lib()
    null = ctypes.c_void_p(0)
    assert L.rr_decode_topk(null, null, null, 1, 1, 4, 4, 1, 0, null, null, null, 0, null) == -1
    num = ctypes.c_int(123)
    rc = L.rr_nms_legacy_host(null, ctypes.cast(ctypes.byref(num), ctypes.c_void_p), null, 0, 5, 0.5, 0)
    assert rc == 0 and num.value == 0          # empty input -> empty keep (nms_wrapper.py:24-25)


def test_ops_fail_loudly_without_cuda():
    import torch
    from rrnet_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    z = torch.zeros(1, 2, 4, 4)
    with pytest.raises(ops.RRNetB200Error):
        ops.decode_topk(z, z, z, 4)
    with pytest.raises(ops.RRNetB200Error):
        ops.focal_forward(z, z)
