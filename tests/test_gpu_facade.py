"""GPU: the drop-in facade.  The reference's own tests (contourist/test/test_tetrahedral.py,
test_triangulated.py) re-expressed against contourist_b200, plus facade-vs-oracle checks."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mp4d, mt2d, mt3d

pytestmark = pytest.mark.gpu


def two_dots3(x, y, z):
    return 1 if (x == y == z == -8 or x == y == z == 0) else -1


def test_isosurface_two_dots(engine):
    """test_tetrahedral.py:13-37.  The reference test is seeded and expects the leak triangles at -9; the full
    scan must contain the 6 in-range expected triangles of the dot at (-8,-8,-8) and the dot at the origin."""
    from contourist_b200 import tetrahedral
    S = tetrahedral.TriangulatedIsosurfaces([-8] * 3, [8] * 3, [2] * 3, two_dots3, 0, [])
    S.search_for_endpoints()
    assert len(S.grid_endpoints) > 0
    (points, triangles) = S.get_points_and_triangles()
    points = [tuple(int(i) for i in pt) for pt in points]
    got = set(frozenset(points[i] for i in triangle) for triangle in triangles)
    for e in ([(-7, -8, -8), (-7, -8, -7), (-7, -7, -7)], [(-8, -8, -7), (-8, -7, -7), (-7, -7, -7)],
              [(-8, -8, -7), (-7, -8, -7), (-7, -7, -7)], [(-8, -7, -8), (-7, -7, -8), (-7, -7, -7)],
              [(-7, -8, -8), (-7, -7, -8), (-7, -7, -7)], [(-8, -7, -8), (-8, -7, -7), (-7, -7, -7)]):
        assert frozenset(e) in got


def test_isosurface_facade_equals_oracle_world_coords(engine):
    from contourist_b200 import tetrahedral

    def f(x, y, z):
        return x * x + y * y + z * z + 0.1 * np.sin(5 * x)
    S = tetrahedral.TriangulatedIsosurfaces([-1] * 3, [1] * 3, [0.125] * 3, f, 0.5, [])
    S.post_process = False                 # the raw engine mesh (the reference's post-processing: tests/test_gpu_post3d.py)
    S.search_for_endpoints()
    pts, tris = S.get_points_and_triangles()
    arr = S.grid.samples(1)
    assert arr.shape == (18, 18, 18)
    r = mt3d.extract(arr, 0.5)
    # canonical form (vertex numbering is an engine choice; the oracle's is by edge key): each triangle
    # as the sorted tuple of its three world positions, the mesh as the sorted list of those
    world = r["pos"] * 0.125 + (-1.0)
    assert len(pts) == len(world) and len(tris) == len(r["tris"])

    def canon(P, T):
        return sorted(tuple(sorted(tuple(P[i]) for i in t)) for t in np.asarray(T))
    assert canon(np.asarray(pts), tris) == canon(world, r["tris"])
    mx, mn, nseg = mt3d.crossing_segments(arr, 0.5, count_only=True)
    assert len(S.grid_endpoints) == nseg
    v0, v1 = S.grid_endpoints[0]
    assert (arr[tuple(v0)] - 0.5) * (arr[tuple(v1)] - 0.5) < 0
    # three.js emitters run on the facade output
    from contourist_b200 import html_demo
    d = json.loads(html_demo.emit_three_json(S))
    assert len(d["faces"]) == 4 * len(tris) and len(d["vertices"]) == 3 * len(pts)
    page = html_demo.grid_html_page(S)
    assert "THREE.BufferGeometry" in page and ("var MESH_FACES = [[%d, %d, %d]" % tuple(int(i) for i in tris[0])) in page


def test_reference_orientation_matches_reference_on_golden_sphere(engine):
    """Second tier (a14): the reference's final mesh of the 13^3 sphere is wound outward; so is the facade's
    with reference_orientation=True, triangle by triangle where both have the triangle."""
    from contourist_b200 import tetrahedral
    g = np.load(os.path.join(GOLDEN, "mt3d_sphere13.npz"))
    G = tetrahedral.Grid3DContour(12, 12, 12, g["field"], float(g["value"]), None)      # None: full scan
    G.post_process = False
    G.reference_orientation = True
    pts, tris = G.get_points_and_triangles()

    def signed(P, T):
        out = {}
        for t in T:
            a, b, c = (np.asarray(P[i]) for i in t)
            key = frozenset(tuple(np.round(P[i], 9)) for i in t)
            if len(key) == 3:
                out[key] = np.sign(np.cross(b - a, c - a) @ ((a + b + c) / 3 - 6.0))
        return out
    mine, ref = signed(pts, tris), signed(g["final_points"], g["final_tris"])
    common = set(mine) & set(ref)
    assert len(common) > 0.7 * len(ref)
    assert all(mine[k] == ref[k] == 1 for k in common)


@pytest.mark.parametrize("geom64", [True, False])
def test_device_orientation_equals_surface_geometry(engine, geom64):
    """ctr_mt3d_orient_reference against the host restatement of surface_geometry.py:52-140 (SurfaceGeometry
    .orient_triangles) on a field with many components: blobs whose high side is inside (kept), blobs whose high side
    is outside (reversed), sheets cut by the volume boundary."""
    from contourist_b200 import engine as E
    from contourist_b200 import surface_geometry
    n = 48
    x, y, z = np.meshgrid(*(np.arange(n, dtype=np.float64),) * 3, indexing="ij")
    rng = np.random.default_rng(5)
    f = np.zeros((n, n, n))
    for _ in range(14):
        c = rng.uniform(4, n - 4, 3)
        r = rng.uniform(2.5, 6.0)
        sgn = rng.choice([-1.0, 1.0])
        f += sgn * np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (r * r))
    f += 0.02 * (x - n / 2)                                     # an open sheet through the box as well
    flags = E.GEOM_F64 if geom64 else 0
    engine.mt3d_run(f, 0.35, flags=flags)
    raw = engine.mt3d_fetch()
    ncomp, nflip = engine.mt3d_orient_reference()
    got = engine.mt3d_fetch()
    assert np.array_equal(got["verts"], raw["verts"])
    geometry = surface_geometry.SurfaceGeometry(raw["verts"], raw["tris"])
    want = geometry.orient_triangles()
    assert sorted(map(tuple, got["tris"].tolist())) == want
    changed = np.any(got["tris"] != raw["tris"], axis=1)
    assert int(changed.sum()) == nflip and 0 < nflip < len(raw["tris"])
    assert ncomp >= 3
    # a second call finds everything oriented already
    ncomp2, nflip2 = engine.mt3d_orient_reference()
    assert (ncomp2, nflip2) == (ncomp, 0)


def test_device_orientation_equals_oracle_dfs(engine):
    """... and against the oracle's restatement of the reference DFS itself (oracle/mt3d.py orient, pinned to the
    reference's final meshes in tests/test_oracle_golden_3d.py::test_orient_matches_reference_final_mesh) on closed blobs of both signs.  Triangles compared up to
    rotation (the DFS starts every triangle at the shared edge)."""
    from contourist_b200 import engine as E
    n = 30
    x, y, z = np.meshgrid(*(np.arange(n, dtype=np.float64),) * 3, indexing="ij")
    rng = np.random.default_rng(11)
    f = np.zeros((n, n, n))
    for c, sgn in zip(([8.3, 8.1, 8.7], [21.2, 9.4, 8.9], [8.8, 21.6, 20.9], [20.7, 20.3, 21.1]), (1, -1, -1, 1)):
        r = rng.uniform(3.0, 4.5)
        f += sgn * np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (r * r))
    for value in (0.4, -0.4):
        engine.mt3d_run(f, value, flags=E.GEOM_F64)
        raw = engine.mt3d_fetch()
        ncomp, nflip = engine.mt3d_orient_reference()
        got = engine.mt3d_fetch()["tris"]
        assert ncomp == 2 and nflip == (len(got) if value > 0 else 0)     # high side inside: engine winds inward

        def rot(t):
            t = tuple(int(i) for i in t)
            k = t.index(min(t))
            return t[k:] + t[:k]
        assert sorted(rot(t) for t in got) == sorted(rot(t) for t in mt3d.orient(raw["verts"], raw["tris"]))


def test_device_orientation_needs_a_run(engine):
    from contourist_b200 import engine as E
    e2 = E.Engine(0)
    with pytest.raises(E.EngineError):
        e2.mt3d_orient_reference()
    e2.close()


def test_grid2d_line_and_dot(engine):
    """test_triangulated.py:79-106."""
    from contourist_b200 import triangulated
    G = triangulated.Grid2DContour(2, 2, lambda x, y: x + y, 1.5, [[(0, 0), (2, 2)]])
    [(closed, contour)] = G.get_contour_sequences()
    assert not closed
    expected = np.array([(1.0, 0.5), (0.75, 0.75), (0.5, 1.0)])
    assert np.allclose(expected, contour) or np.allclose(expected, contour[::-1])

    def dot(x, y):
        return 2 if (x == 1 and y == 1) else 0
    G = triangulated.Grid2DContour(3, 3, dot, 1, [[(0, 0), (1, 1)]])
    [(closed, contour)] = G.get_contour_sequences()
    assert closed
    expected = [[0.5, 0.5], [1.0, 0.5], [1.5, 1.0], [1.5, 1.5], [1.0, 1.5], [0.5, 1.0]]
    rots = [expected[s:] + expected[:s] for s in range(6)]
    rots += [list(reversed(x)) for x in rots]
    assert any(np.allclose(np.array(x), contour) for x in rots)


def test_dxdy_two_dots(engine):
    """test_triangulated.py:44-68 (broken at reference HEAD by a missing import; expected values are its own)."""
    from contourist_b200 import triangulated

    def two_dots(x, y):
        return 1 if (x == y == -4 or x == y == 0) else -1
    C = triangulated.DxDy2DContour(-4, -4, 4, 4, 2, 2, two_dots, 0)
    contours = [(c, [(int(x * 10), int(y * 10)) for (x, y) in pts]) for c, pts in C.get_contour_sequences()]
    exp_open = [(-40, -30), (-30, -30), (-30, -40)]
    exp_closed = [(0, 10), (10, 10), (10, 0), (0, -10), (-10, -10), (-10, 0)]
    assert len(contours) == 2
    (c0, p0), (c1, p1) = sorted(contours, key=lambda x: x[0])
    assert not c0 and (p0 == exp_open or p0 == exp_open[::-1])
    rots = [exp_closed[s:] + exp_closed[:s] for s in range(6)]
    rots += [list(reversed(x)) for x in rots]
    assert c1 and p1 in rots


def test_multiple_2d_contour_dictionary(engine):
    from contourist_b200 import multiple_2d_contour

    def f(x, y):
        return np.sqrt(np.sin(3 * x + y * y) ** 2 + np.cos(4 * y + x * x) ** 2)
    C = multiple_2d_contour.Linear2DContour(-2, -2, 2, 2, 0.125, 0.125, f, breakpoints=5)
    arr = C.grid.samples(0)
    assert np.allclose(C.values, mt2d.linear_levels(arr, 5))
    D = C.get_contours_dictionary()
    assert sorted(D) == sorted(C.values)
    for value in C.values:
        r = mt2d.extract_level(arr, value)
        mine = D[value]
        ref = mt2d.polylines(r["keys"], r["pos"], r["seg_keys"])
        assert len(mine) == len(ref)
        assert sorted(c for c, _ in mine) == sorted(c for c, _, _ in ref)
        allp = np.concatenate([np.asarray(p) for _, p in mine])
        refp = np.concatenate([p for _, _, p in ref]) * 0.125 - 2.0
        assert np.array_equal(np.unique(allp, axis=0), np.unique(refp, axis=0))
    P = multiple_2d_contour.Percentile2DContour(-2, -2, 2, 2, 0.25, 0.25, f, breakpoints=4)
    assert np.allclose(P.values, mt2d.percentile_levels(P.grid.samples(0), 4))


def test_morphing_isosurfaces_json(engine):
    from contourist_b200 import pentatopes

    def fg(x, y, z, t):
        return t * 3 * np.sqrt(x * x + z * z) + (1 - t) * 3 * np.sqrt((1 - np.sqrt(x * x + y * y)) ** 2 + z * z)
    G = pentatopes.MorphingIsoSurfaces([-2, -2, -2, 0], [2, 2, 2, 1], [0.5, 0.5, 0.5, 0.25], fg, 1.2, [])
    G.search_for_endpoints()
    mt = G.collect_morph_triangles()
    arr = G.grid.samples(1)
    r = mp4d.extract(arr, 1.2)
    corner = np.array(arr.shape) - 1
    bp = mp4d.bin_times(r["pos"], corner[3])
    keep = mp4d.drop_instant(bp, r["tets"]) & ~mp4d.tiny_mask(bp, r["tets"], corner)
    segs, tris = mp4d.morph_triangles(bp, r["tets"][keep])
    assert np.array_equal(mt.points4d, bp * G.grid.delta + G.grid.mins)
    assert np.array_equal(np.sort(mt.segment_point_indices, axis=1), np.sort(segs, axis=1))
    assert np.array_equal(np.sort(mt.triangle_segment_indices, axis=1)[np.lexsort(np.sort(mt.triangle_segment_indices, axis=1).T[::-1])],
                          np.sort(tris, axis=1)[np.lexsort(np.sort(tris, axis=1).T[::-1])])
    d = json.loads(G.to_json())
    assert d["counts"] == [len(bp), len(segs), len(tris)]
    assert len(d["positions"]) == 4 * len(bp) and max(d["positions"]) <= 999999 and min(d["positions"]) >= 0
    assert d["min_value"] >= 0.0 and d["max_value"] <= 1.0
