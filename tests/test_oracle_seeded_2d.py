"""CPU: the oracle's 2D seeded tracking (oracle/mt2d.py seeded_keys) against runs of the unmodified reference tracker
with explicit seed segments (tests/golden/make_golden.py seeded2d)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mt2d

FILES = sorted(glob.glob(os.path.join(GOLDEN, "seeded2d_*.npz")))


def golden_keys(g):
    n1 = g["field"].shape[1]
    pm = np.minimum(g["low"], g["high"])
    d = np.maximum(g["low"], g["high"]) - pm
    lowmin = (g["low"] == pm).all(axis=1)
    return np.sort(mt2d.key_of(pm[:, 0] * n1 + pm[:, 1], 0, lowmin) | ((d[:, 0] * 2 + d[:, 1]).astype(np.uint64) << np.uint64(1)))


def test_have_seeded2d_goldens():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_seeded_pairs_match_reference(path):
    g = np.load(path)
    field, z = g["field"], float(g["value"])
    got = mt2d.seeded_keys(field, z, g["seeds"])
    assert np.array_equal(got, golden_keys(g))
    full = mt2d.extract_level(field, z)
    assert len(got) < len(full["keys"])                       # contours the seeds do not reach are left out
    r = mt2d.extract_level_seeded(field, z, g["seeds"])
    lines = mt2d.polylines(r["keys"], r["pos"], r["seg_keys"])
    assert sorted(len(p) for _, _, p in lines) == sorted(g["length"].tolist())
    assert sorted(bool(c) for c, _, _ in lines) == sorted(bool(c) for c in g["closed"])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_facade_seed_filter_matches_reference(path):
    """The facade's host-side filter (triangulated.seeded_segments) applied to a full segment soup gives the pairs the
    reference tracker found."""
    from contourist_b200 import triangulated
    g = np.load(path)
    f, z = g["field"], float(g["value"])
    r = mt2d.extract_level(f, z)
    keep = triangulated.seeded_segments(lambda p: float(f[int(p[0]), int(p[1])]), z, f.shape[1], r["seg_keys"], g["seeds"])
    assert np.array_equal(np.unique(r["seg_keys"][keep]), golden_keys(g))
