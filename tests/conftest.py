import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Built on demand with gcc."""
    import oracle
    oracle.build()
    return oracle


def rel_err(a, b, floor=1e-6):
    """max |a-b| / max(|b|, floor-scaled magnitude) -- the 1e-5 relative fp32 contract."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), floor * max(1.0, float(np.abs(b).max()) if b.size else 1.0))
    return float(np.max(np.abs(a - b) / scale))


def box_rel_err(a, b):
    """Error of [n,>=4] box rows relative to each row's own scale max(|x|,|y|,|w|,|h|): coordinates come
    from sums that cancel (x = ctr - w/2), so a per-element relative error is not meaningful."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b[:, :4]).max(axis=1, keepdims=True), 1.0)
    e = np.abs(a - b)
    e[:, :4] /= scale
    e[:, 4:] /= np.maximum(np.abs(b[:, 4:]), 1e-6)
    return float(e.max())
