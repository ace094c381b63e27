"""CPU: the oracle's seeded tracking (oracle/mt3d.py initial_voxels / flood_fill / extract_seeded) against runs of
the unmodified reference tracker with explicit seed segments (tests/golden/make_golden.py seeded), and the facade's
host-side restatement of find_initial_voxels against both."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mt3d

FILES = sorted(glob.glob(os.path.join(GOLDEN, "seeded3d_*.npz")))


def test_have_seeded_goldens():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_seeded_tracking_matches_reference(path):
    g = np.load(path)
    field, value, seeds = g["field"], float(g["value"]), g["seeds"]
    init = mt3d.initial_voxels(field, value, seeds)
    assert sorted(init) == [tuple(v) for v in g["initial"].tolist()]
    mask = mt3d.flood_fill(field, value, init)
    assert np.array_equal(np.argwhere(mask), g["voxels"])                 # both sorted lexicographically
    r = mt3d.extract_seeded(field, value, seeds)
    assert (len(r["keys"]), len(r["tris"])) == (int(g["n_keys"]), int(g["n_tris"]))
    n1, n2 = field.shape[1:]
    pmin = np.minimum(g["key_low"], g["key_high"])
    d = np.maximum(g["key_low"], g["key_high"]) - pmin
    gk = ((((pmin[:, 0] * n1 + pmin[:, 1]) * n2 + pmin[:, 2]).astype(np.uint64) << np.uint64(3))
          | (d[:, 0] * 4 + d[:, 1] * 2 + d[:, 2]).astype(np.uint64))
    assert np.array_equal(np.sort(gk), r["keys"])
    full = mt3d.extract(field, value)
    if "wave_sheet" in path:
        assert len(r["tris"]) == len(full["tris"])                        # one sheet: the tracker finds all of it
    else:
        assert len(r["tris"]) < len(full["tris"])                         # components the seeds do not reach are left out


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_facade_initial_voxels_match_reference(path):
    from contourist_b200 import tetrahedral
    g = np.load(path)
    got = tetrahedral.initial_voxels(g["field"], float(g["value"]), g["seeds"])
    assert got.tolist() == g["initial"].tolist()


def test_reference_known_answer_two_dots_seeded():
    """contourist/test/test_tetrahedral.py:13-37: seeded on the dot at (-8,-8,-8), the dot at the origin is NOT
    returned; the 6 in-range triangles are (the other 2 of the reference's 8 hang on out-of-range leak voxels)."""
    def two_dots(x, y, z):
        return 1 if (x == y == z == -8 or x == y == z == 0) else -1
    field = np.array([[[two_dots(-8 + 2 * i, -8 + 2 * j, -8 + 2 * k) for k in range(10)] for j in range(10)]
                      for i in range(10)], dtype=np.float64)
    r = mt3d.extract_seeded(field, 0.0, [[(0, 0, 0), (0, 0, 8)]])
    world = mt3d.to_world(r["pos"], [-8.0] * 3, [2.0] * 3)
    pts = [tuple(int(c) for c in p) for p in world]
    got = set(frozenset(pts[i] for i in t) for t in r["tris"])
    assert got == {frozenset([(-7, -8, -8), (-7, -8, -7), (-7, -7, -7)]), frozenset([(-8, -8, -7), (-8, -7, -7), (-7, -7, -7)]),
                   frozenset([(-8, -8, -7), (-7, -8, -7), (-7, -7, -7)]), frozenset([(-8, -7, -8), (-7, -7, -8), (-7, -7, -7)]),
                   frozenset([(-7, -8, -8), (-7, -7, -8), (-7, -7, -7)]), frozenset([(-8, -7, -8), (-8, -7, -7), (-7, -7, -7)])}
