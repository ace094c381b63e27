"""CPU: the numpy oracle (oracle/mt3d.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's own known-answer test."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, c1_golden
from oracle import mt3d

FILES = sorted(glob.glob(os.path.join(GOLDEN, "mt3d_*.npz")))


def golden_keys(g):
    n0, n1, n2 = g["field"].shape
    klow, khigh = g["key_low"], g["key_high"]
    pmin = np.minimum(klow, khigh)
    d = np.maximum(klow, khigh) - pmin
    keys = ((((pmin[:, 0] * n1 + pmin[:, 1]) * n2 + pmin[:, 2]).astype(np.uint64) << np.uint64(3))
            | (d[:, 0] * 4 + d[:, 1] * 2 + d[:, 2]).astype(np.uint64))
    lowmin = (klow == pmin).all(axis=1).astype(np.uint8)
    return keys, lowmin


def test_have_goldens():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_raw_extraction_matches_reference(path):
    check_raw_extraction(np.load(path))


def test_c1_sphere65_matches_full_reference_run():
    """BASELINE configs[0]: the one config the whole reference runs; SURVEY.md 8(d) C1 quotes the same counts."""
    g = c1_golden()
    assert (len(g["voxels"]), len(g["key_low"]), len(g["tris"])) == (9656, 28778, 57552)
    assert (int(g["n_final_points"]), int(g["n_final_tris"])) == (28730, 57456)
    check_raw_extraction(g)
    r = mt3d.extract(g["field"], 0.5)
    crossing_tets = sum(int(mt3d.tet_emits((r["codes"] >> np.uint32(5 * k)) & np.uint32(31)).sum()) for k in range(6))
    assert crossing_tets == 43956


def check_raw_extraction(g):
    field, value = g["field"], float(g["value"])
    n0, n1, n2 = field.shape
    r = mt3d.extract(field, value)
    # P-active: emitting voxels  <=  reference's surface voxels  <=  border_voxel predicate
    gv = g["voxels"]
    gvl = (gv[:, 0] * (n1 - 1) + gv[:, 1]) * (n2 - 1) + gv[:, 2]
    act = np.nonzero(mt3d.active_cells(field, value).reshape(-1))[0]
    assert np.isin(r["cells"], gvl).all()
    assert np.isin(gvl, act).all()
    # P-keys, key orientation, P-pos (0 ulp)
    gkeys, glowmin = golden_keys(g)
    o = np.argsort(gkeys)
    assert np.array_equal(gkeys[o], r["keys"])
    assert np.array_equal(glowmin[o], r["lowmin"])
    assert np.array_equal(g["key_pos"][o], r["pos"])
    # voxels that own a simplex in the reference == emitting voxels
    owner = np.minimum(g["key_low"], g["key_high"])[g["tris"]].min(axis=1)
    own_lin = np.unique((owner[:, 0] * (n1 - 1) + owner[:, 1]) * (n2 - 1) + owner[:, 2])
    assert np.array_equal(own_lin, r["cells"])
    # P-topo: triangles equal modulo the diagonal of 2-2 quads (reference: CPython set order)
    gt = set(map(tuple, np.sort(gkeys[g["tris"]], axis=1).tolist()))
    assert len(gt) == len(r["tri_keys"])
    tk = r["tri_keys"]
    cell, tet = r["tri_cell"], r["tri_tet"]
    i = 0
    n_quads = 0
    while i < len(tk):
        if i + 1 < len(tk) and cell[i + 1] == cell[i] and tet[i + 1] == tet[i]:
            ad, ac, bc = (int(x) for x in tk[i])
            ad2, bd, bc2 = (int(x) for x in tk[i + 1])
            assert (ad, bc) == (ad2, bc2)
            split1 = {tuple(sorted((ad, ac, bc))), tuple(sorted((ad, bd, bc)))}
            split2 = {tuple(sorted((ac, ad, bd))), tuple(sorted((ac, bc, bd)))}
            assert split1 <= gt or split2 <= gt
            n_quads += 1
            i += 2
        else:
            assert tuple(sorted(int(x) for x in tk[i])) in gt
            i += 1
    assert n_quads > 0


def test_plateau_exercises_allclose_skip():
    g = np.load(os.path.join(GOLDEN, "mt3d_plateau6.npz"))
    code = mt3d.tet_cases(g["field"], float(g["value"]))
    skipped = 0
    for k in range(6):
        c5 = (code >> np.uint32(5 * k)) & np.uint32(31)
        m = c5 & 15
        skipped += int(((c5 & 16) != 0).__and__(m != 0).__and__(m != 15).sum())
    assert skipped > 0, "fixture must contain crossing tets skipped by np.allclose"


def test_crossing_segments_match_reference_seed_count():
    for path in FILES:
        g = np.load(path)
        _, _, n = mt3d.crossing_segments(g["field"], float(g["value"]), count_only=True)
        assert n == int(g["n_seeds"])


def test_reference_known_answer_two_dots():
    """contourist/test/test_tetrahedral.py:13-37.  The reference test is seeded (only the dot at (-8,-8,-8)
    is found) and includes vertices at -9 from out-of-range 'leak' voxels (hard part 2); the
    in-range expected triangles must all be produced by the full scan."""
    def two_dots(x, y, z):
        return 1 if (x == y == z == -8 or x == y == z == 0) else -1
    mins, delta = -8.0, 2.0
    n = 10                                  # samples 0..9 <-> N = 9 = int(16/2)+1 grid dimensions + 1
    field = np.array([[[two_dots(mins + i * delta, mins + j * delta, mins + k * delta) for k in range(n)]
                       for j in range(n)] for i in range(n)], dtype=np.float64)
    r = mt3d.extract(field, 0.0)
    world = mt3d.to_world(r["pos"], [mins] * 3, [delta] * 3)
    pts = [tuple(int(c) for c in p) for p in world]
    got = set(frozenset(pts[i] for i in t) for t in r["tris"])
    expected = [frozenset([(-7, -8, -8), (-7, -8, -7), (-7, -7, -7)]),
                frozenset([(-8, -8, -7), (-8, -7, -7), (-7, -7, -7)]),
                frozenset([(-8, -8, -7), (-7, -8, -7), (-7, -7, -7)]),
                frozenset([(-8, -7, -8), (-7, -7, -8), (-7, -7, -7)]),
                frozenset([(-7, -8, -8), (-7, -7, -8), (-7, -7, -7)]),
                frozenset([(-8, -7, -8), (-8, -7, -7), (-7, -7, -7)])]
    for e in expected:
        assert e in got


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_orient_matches_reference_final_mesh(path):
    """oracle.orient (surface_geometry.py:52-140) on the reference's own final mesh must give back the reference's
    oriented triangles (its last step is exactly this call).  Compared up to rotation of each triple.  ints7 (a
    non-manifold integer field full of ties) is only required to agree as unoriented triangles: there the
    reference's DFS depends on CPython set order."""
    g = np.load(path)
    fp, ft = g["final_points"], g["final_tris"]

    def rot(t):
        t = tuple(int(i) for i in t)
        k = t.index(min(t))
        return t[k:] + t[:k]
    got = sorted(rot(t) for t in mt3d.orient(fp, ft))
    want = sorted(rot(t) for t in ft)
    if "ints7" in path:
        assert sorted(tuple(sorted(t)) for t in got) == sorted(tuple(sorted(t)) for t in want)
    else:
        assert got == want
