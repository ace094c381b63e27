import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C ABI.  No fallback: a missing library or device is an error."""
    from contourist_b200 import engine as E
    return E.default_engine(0)


def golden_path(name):
    return os.path.join(GOLDEN, name)


def c1_golden():
    """BASELINE configs[0]: the reference's own run of f = x^2+y^2+z^2 on [-1,1]^3, delta 1/32
    (N = 65 voxels per axis, samples 0..65), value 0.5 -- tests/golden/make_golden.py c1.  The field is not stored:
    it is this formula."""
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN, "c1_sphere65.npz")))
    x = -1.0 + np.arange(66) / 32.0
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    g["field"] = X * X + Y * Y + Z * Z
    for k in ("voxels", "key_low", "key_high", "tris"):
        g[k] = g[k].astype(np.int64)
    return g
