import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def max_rel(y, ref):
    """Per-channel max-norm relative error: max_t|y-ref| / max_t|ref|,
    worst channel.  NaN/inf positions must coincide and are excluded."""
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    bad_ref = ~np.isfinite(ref)
    assert np.array_equal(bad_ref, ~np.isfinite(y)), "non-finite pattern differs"
    if ref.ndim == 1:
        y, ref, bad_ref = y[None], ref[None], bad_ref[None]
    d = np.where(bad_ref, 0.0, np.abs(y - np.where(bad_ref, 0.0, ref)))
    scale = np.max(np.where(bad_ref, 0.0, np.abs(ref)), axis=-1)
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.max(d, axis=-1) / scale))
