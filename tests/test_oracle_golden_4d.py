"""CPU: the 4D numpy oracle (oracle/mp4d.py) against golden vectors produced by the unmodified reference."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mp4d

FILES = sorted(glob.glob(os.path.join(GOLDEN, "mp4d_*.npz")))


def golden_keys(g):
    n = g["field"].shape
    klow, khigh = g["key_low"], g["key_high"]
    pm = np.minimum(klow, khigh)
    d = np.maximum(klow, khigh) - pm
    gk = ((((((pm[:, 0] * n[1] + pm[:, 1]) * n[2] + pm[:, 2]) * n[3] + pm[:, 3]).astype(np.uint64)) << np.uint64(4))
          | (d[:, 0] * 8 + d[:, 1] * 4 + d[:, 2] * 2 + d[:, 3]).astype(np.uint64))
    return gk, (klow == pm).all(axis=1).astype(np.uint8)


def tri_sets(segs, tris):
    s = np.sort(segs, axis=1)
    return set(frozenset(tuple(int(v) for v in s[x]) for x in t) for t in tris)


def check_triangles_cover_polygons(v4, tets, triangles, t_eps):
    """The reference splits a 4-edge slice along `interpolated[0]` + its disjoint edge, which depends on ITS
    vertex numbering (morph_geometry.py:159,171-186).  Split-independent statement: every slice polygon of
    every tetrahedron is covered by triangles of the reference (1 for a triangle, 2 sharing two disjoint
    edges for a quad), and every reference triangle lies in some polygon."""
    polys = set()
    for tet in set(frozenset(int(x) for x in t) for t in tets):
        if len(tet) == 4:
            for inter in mp4d.slice_polygons(v4, tet):
                polys.add(frozenset(inter))

    def dead(tri):
        return any(abs(v4[i][-1] - v4[j][-1]) <= t_eps for (i, j) in tri)
    expected = 0
    for poly in polys:
        inside = [t for t in triangles if t <= poly]
        if len(poly) == 3:
            if not dead(poly):
                assert poly in triangles
        elif len(poly) == 4:
            # both possible splits, minus triangles with a zero-duration segment
            assert len(inside) >= 1 or all(dead(frozenset(c)) for c in __import__("itertools").combinations(poly, 3))
    for t in triangles:
        assert any(t <= poly for poly in polys)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_raw_extraction_and_morph_stage_match_reference(path):
    g = np.load(path)
    field, value = g["field"], float(g["value"])
    n = field.shape
    r = mp4d.extract(field, value)
    gk, glow = golden_keys(g)
    o = np.argsort(gk)
    inv = np.empty_like(o)
    inv[o] = np.arange(len(o))
    assert np.array_equal(gk[o], r["keys"])
    assert np.array_equal(glow[o], r["lowmin"])
    assert np.array_equal(g["key_pos"][o], r["pos"])
    # hypervoxels that own a tetrahedron
    gv = g["voxels"]
    cn = [s - 1 for s in n]
    gvl = ((gv[:, 0] * cn[1] + gv[:, 1]) * cn[2] + gv[:, 2]) * cn[3] + gv[:, 3]
    assert np.isin(r["cells"], gvl).all()
    # tetrahedra: same count; identical for 1-vs-4 pentatopes; 2-vs-3 prisms cover the same 6 keys
    gt = set(map(tuple, np.sort(inv[g["tets"]], axis=1).tolist()))
    assert len(gt) == len(r["tets"])
    tk, cell, pent = np.sort(r["tets"], axis=1), r["tet_cell"], r["tet_pent"]
    gold_by_keyset = {}
    i = 0
    singles = prisms = 0
    gold_union = {}
    for t in gt:
        for k in t:
            gold_union.setdefault(k, set()).add(t)
    while i < len(tk):
        if i + 2 < len(tk) and cell[i + 2] == cell[i] and pent[i + 2] == pent[i]:
            six = set(tk[i]) | set(tk[i + 1]) | set(tk[i + 2])
            assert len(six) == 6
            # the reference has exactly 3 tets inside this 6-key prism
            inside = set(t for k in six for t in gold_union[k] if set(t) <= six)
            assert len(inside) >= 3
            prisms += 1
            i += 3
        else:
            assert tuple(tk[i]) in gt
            singles += 1
            i += 1
    assert singles > 0 and prisms > 0
    # bin_times, drop_instant (on the reference's own tets, renumbered)
    corner = np.array(n) - 1
    bp = mp4d.bin_times(r["pos"], corner[3])
    assert np.array_equal(bp, g["binned_pos"][o])
    gtets = inv[g["tets"]]
    keep = mp4d.drop_instant(bp, gtets)
    assert (set(map(tuple, np.sort(gtets[keep], axis=1).tolist()))
            == set(map(tuple, np.sort(inv[g["tets_after_drop"]], axis=1).tolist())))
    tiny = mp4d.tiny_mask(bp, gtets[keep], corner)
    assert len(g["tets_after_tiny"]) == int((~tiny).sum())
    # slicing
    segs, tris = mp4d.morph_triangles(bp, gtets[keep][~tiny])
    gs = set(map(tuple, np.sort(inv[g["morph_segments"]], axis=1).tolist()))
    assert gs == set(map(tuple, np.sort(segs, axis=1).tolist()))
    assert len(tris) == len(g["morph_triangles"])
    tv = bp[:, 3]
    t_eps = 1e-7 * (tv.max() - tv.min())
    check_triangles_cover_polygons(bp, gtets[keep][~tiny], tri_sets(inv[g["morph_segments"]], g["morph_triangles"]), t_eps)
    check_triangles_cover_polygons(bp, gtets[keep][~tiny], tri_sets(segs, tris), t_eps)
    # segments are stored low-t first (morph_geometry.py:13-17)
    assert (bp[segs[:, 0], 3] <= bp[segs[:, 1], 3]).all()


def test_instant_tets_are_exercised():
    g = np.load(os.path.join(GOLDEN, "mp4d_ints4.npz"))
    assert len(g["tets_after_drop"]) < len(g["tets"])
