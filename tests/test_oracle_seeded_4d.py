"""CPU: the 4D seeded tracking of oracle/mp4d.py against runs of the unmodified reference's GridContour4D with explicit
seed segments (tests/golden/make_golden.py seeded4d; pentatopes.py:92-106, tetrahedral.py:396-469 with OFFSETS4D)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mp4d

FILES = sorted(glob.glob(os.path.join(GOLDEN, "seeded4d_*.npz")))


def golden_keys4(g):
    n = g["field"].shape
    klow, khigh = g["key_low"].astype(np.int64), g["key_high"].astype(np.int64)
    pmin = np.minimum(klow, khigh)
    d = np.maximum(klow, khigh) - pmin
    lin = ((pmin[:, 0] * n[1] + pmin[:, 1]) * n[2] + pmin[:, 2]) * n[3] + pmin[:, 3]
    return (lin.astype(np.uint64) << np.uint64(4)) | (d[:, 0] * 8 + d[:, 1] * 4 + d[:, 2] * 2 + d[:, 3]).astype(np.uint64)


def test_have_goldens():
    assert len(FILES) == 3


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_seeded_4d_matches_reference_tracker(path):
    g = np.load(path)
    field, value, seeds = g["field"], float(g["value"]), g["seeds"]
    start = mp4d.initial_voxels(field, value, seeds)
    assert sorted(start) == [tuple(v) for v in g["initial"].tolist()]
    mask = mp4d.flood_fill(field, value, start)
    assert np.array_equal(np.argwhere(mask), g["voxels"])
    r = mp4d.extract_seeded(field, value, seeds)
    assert np.array_equal(r["keys"], np.sort(golden_keys4(g)))
    assert len(r["tets"]) == int(g["n_tets"])
    full = mp4d.extract(field, value)
    assert (len(r["tets"]) < len(full["tets"])) == ("both" not in path)
