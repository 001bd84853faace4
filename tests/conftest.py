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
