"""CPU: the reference's own host-side tests re-expressed against the drop-in facade
(contourist/test/test_field2d.py, test_triangulated.py TestMisc / TestAdjacentPairs), plus facade host logic."""
import numpy as np
import pytest

from contourist_b200 import field2d, grid_field, morph_geometry, surface_geometry, triangulated

EXPECT_SVG = """
<svg height="300.0" width="300" viewBox="-1 -1 2 2">
<path stroke-width="0.02" stroke="black" fill="none" d="M0.00 0.00 L0.00 1.00 L1.00 1.00 Z" />
<path stroke-width="0.02" stroke="black" fill="none" d="M-1.00 -1.00 L-1.00 0.00" />
</svg>
"""


def make_grid(materialize, cache):
    def function(x, y):
        return (x + 100) * 1000 + (y + 100)
    return field2d.Function2DGrid(-10, -20, 30, 50, 10.0, 20.0, function, materialize, cache)


@pytest.mark.parametrize("materialize,cache", [(False, False), (False, True), (True, False)])
def test_f2dgrid(materialize, cache):
    grid = make_grid(materialize, cache)
    for _ in (1, 2):
        assert np.allclose(grid.to_grid_coordinates((-10, -20)), (0, 0))
        assert np.allclose(grid.from_grid_coordinates((0, 0)), (-10, -20))
        assert np.allclose(grid.to_grid_coordinates((0, 0)), (1, 1))
        assert np.allclose(grid.from_grid_coordinates((1, 1)), (0, 0))
        assert np.allclose(grid.grid_function(0, 0), 90080)
        assert np.allclose(grid.grid_function(4, 3), 130140)
        S = set(tuple(int(v) for v in x) for x in grid.surrounding_vertices((5, 5)))
        assert S == set([(1, 2), (1, 1), (2, 1), (2, 2)])
    if materialize:
        expect = [[90080.0, 90100.0, 90120.0, 90140.0], [100080.0, 100100.0, 100120.0, 100140.0],
                  [110080.0, 110100.0, 110120.0, 110140.0], [120080.0, 120100.0, 120120.0, 120140.0],
                  [130080.0, 130100.0, 130120.0, 130140.0]]
        assert np.allclose(grid.materialized_array, expect)
        assert grid.cache == {}
    elif cache:
        assert grid.materialized_array is None
        assert grid.cache == {(0, 0): 90080.0, (4, 3): 130140.0}
    else:
        assert grid.materialized_array is None and len(grid.cache) == 0


def test_samples_vectorised_equals_loop():
    import math

    def vec(x, y, z):
        return x * x + np.sin(y) - z

    def scalar(x, y, z):
        return x * x + math.sin(y) - z          # math.sin rejects arrays -> per-point loop
    a = grid_field.FunctionGrid([0, 0, 0], [1, 2, 1], [0.25, 0.5, 0.5], vec).samples(1)
    b = grid_field.FunctionGrid([0, 0, 0], [1, 2, 1], [0.25, 0.5, 0.5], scalar).samples(1)
    assert a.shape == (6, 6, 4) and np.allclose(a, b)


def test_svg():
    cseqs = [(True, [(0, 0), (0, 1), (1, 1)]), (False, [(-1, -1), (-1, 0)])]
    assert triangulated.contour_sequences_to_svg(cseqs).strip() == EXPECT_SVG.strip()


def test_adjacent_pairs():
    adj = list(triangulated.adjacent_pairs((0, 0), (0, 1)))
    assert adj == [((0, 0), (-1, 0)), ((0, 0), (1, 1)), ((1, 1), (0, 1)), ((-1, 0), (0, 1))]
    assert set(triangulated.adjacent_pairs((0, 0), (1, 0))) == set(
        [((0, -1), (1, 0)), ((0, 0), (0, -1)), ((1, 1), (1, 0)), ((0, 0), (1, 1))])
    assert set(triangulated.adjacent_pairs((0, 0), (-1, -1))) == set(
        [((0, -1), (-1, -1)), ((-1, 0), (-1, -1)), ((0, 0), (0, -1)), ((0, 0), (-1, 0))])


def test_chain_segments_open_and_closed():
    k = np.array([[1, 2], [2, 3], [10, 11], [11, 12], [12, 10]], dtype=np.uint64)
    pos = {1: (0, 0), 2: (1, 0), 3: (2, 0), 10: (5, 5), 11: (6, 5), 12: (5, 6)}
    p = np.array([[pos[int(a)], pos[int(b)]] for a, b in k], dtype=float)
    out = triangulated.chain_segments(k, p)
    assert [(c, len(q)) for c, q in out] == [(False, 3), (True, 3)]
    assert np.allclose(out[0][1], [(0, 0), (1, 0), (2, 0)])


def test_orient_triangles_outward_two_components():
    V = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)
    T = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]])
    rng = np.random.default_rng(0)
    V2 = np.concatenate([V, V * 0.5 + 5.0])
    T2 = np.concatenate([T, T + 6])
    T2 = np.array([t[::-1] if rng.random() < 0.5 else t for t in T2])
    o = np.array(surface_geometry.SurfaceGeometry(V2, T2).orient_triangles())
    n = np.cross(V2[o[:, 1]] - V2[o[:, 0]], V2[o[:, 2]] - V2[o[:, 0]])
    cen = V2[o].mean(axis=1)
    cen = np.where((cen > 2.5).any(axis=1)[:, None], cen - 5.0, cen)
    assert ((n * cen).sum(axis=1) > 0).all()


def test_morph_triangles_json_layout():
    pts = [(0, 0, 0, 0), (0, 0, 1, 0), (2, 3, 2, 3), (3, 2, 3, 5)]
    mt = morph_geometry.MorphTriangles(pts, [(0, 1), (2, 1), (0, 2), (1, 3)], [(0, 1, 2), (0, 2, 3)])
    assert mt.segment_point_indices.tolist() == [[0, 1], [1, 2], [0, 2], [1, 3]]       # low t first
    import json
    d = json.loads(mt.to_json())
    assert d["counts"] == [4, 4, 2] and d["max_value"] == 5.0 and d["min_value"] == 0.0
    assert d["positions"][-4:] == [999999, 666666, 999999, 999999]
    assert d["segments"] == [0, 1, 1, 2, 0, 2, 1, 3] and len(d["triangles"]) == 6


def test_engine_import_fails_loudly_without_library(monkeypatch, tmp_path):
    from contourist_b200 import engine
    monkeypatch.setattr(engine, "_lib", None)
    monkeypatch.setattr(engine, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(engine.EngineError):
        engine.load_library()


def test_canonical_mesh_sorts_by_key_and_remaps_triangles():
    """engine.canonical_mesh: vertices by edge key, triangle ids remapped, ids of the next shard left alone."""
    from contourist_b200 import engine as E
    keys = np.array([40, 8, 24, 16], dtype=np.uint64)
    out = dict(keys=keys, lowmin=np.array([1, 0, 1, 0], dtype=np.uint8),
               verts=np.arange(12, dtype=np.float64).reshape(4, 3), normals=None,
               tris=np.array([[0, 1, 2], [3, 2, 5]], dtype=np.int32))
    c = E.canonical_mesh(out)
    assert c["keys"].tolist() == [8, 16, 24, 40]
    assert c["lowmin"].tolist() == [0, 0, 1, 1]
    assert np.array_equal(c["verts"], out["verts"][[1, 3, 2, 0]])
    # old id -> new id: 0->3, 1->0, 2->2, 3->1; id 5 >= n_verts belongs to the next shard and stays
    assert c["tris"].tolist() == [[3, 0, 2], [1, 2, 5]]
    assert out["tris"].tolist() == [[0, 1, 2], [3, 2, 5]]          # input untouched


@pytest.mark.parametrize("name", ["sphere13", "wave11", "noise8", "ints7", "plateau6"])
def test_clean_triangles_against_reference(name):
    """surface_geometry.clean_triangles (surface_geometry.py:14-50) on the raw mesh of a 3D golden, against the
    unmodified reference's result on the same input (tests/golden/make_golden.py clean): the same triangles as sets
    of positions.  Vertex counts may differ -- the reference merges coincident vertices only in the order its set of
    triangles happens to be iterated, the facade merges all of them.  plateau6 (samples within 1e-7 of the isovalue):
    the reference keeps a few triangles between vertices that a LATER zero-area triangle would have merged; the
    facade's are a subset."""
    import os
    from conftest import GOLDEN
    from contourist_b200 import surface_geometry
    raw = np.load(os.path.join(GOLDEN, "mt3d_%s.npz" % name))
    ref = np.load(os.path.join(GOLDEN, "clean3d_%s.npz" % name))
    geometry = surface_geometry.SurfaceGeometry(raw["key_pos"], raw["tris"])
    verts, tris = geometry.clean_triangles()

    digits = 3 if name == "plateau6" else 7            # merged vertices keep different representatives (allclose apart)

    def geo(P, T):
        P = np.asarray(P, dtype=float)
        return set(tuple(sorted(tuple(np.round(P[i], digits)) for i in t)) for t in np.asarray(T))
    mine, theirs = geo(verts, tris), geo(ref["vertices"], ref["triangles"])
    if name == "plateau6":
        assert mine <= theirs and len(theirs) - len(mine) <= 12
    else:
        assert mine == theirs
    assert len(verts) <= len(ref["vertices"])
    assert all(len(set(t)) == 3 for t in np.asarray(tris).tolist())


def test_morph_triangles_to_json_bytes_equal_reference():
    """MorphTriangles.to_json (morph_geometry.py:91-128; number lists through ctr_wire_format) on the morph arrays of
    a 4D golden: byte for byte the text the unmodified reference produced (tests/golden/make_golden.py json4d)."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    want = json.load(open(os.path.join(GOLDEN, "mp4d_to_json.json")))["morph7"]
    g = np.load(os.path.join(GOLDEN, "mp4d_morph7.npz"))
    text = morph_geometry.MorphTriangles(g["morph_points"], g["morph_segments"], g["morph_triangles"]).to_json()
    assert text[:200] == want["head"] and len(text) == want["length"]
    assert hashlib.sha256(text.encode()).hexdigest() == want["sha256"]


def test_svg_text_equals_reference():
    """contour_sequences_to_svg (triangulated.py:16-50): the text the unmodified reference produced for this input."""
    rng = np.random.default_rng(0)
    seqs = [(True, [np.array(p) for p in rng.uniform(-3, 7, (9, 2))]), (False, [np.array(p) for p in rng.uniform(0, 1, (4, 2))])]
    want = ('\n<svg height="340.1828516632472" width="300" viewBox="-2.5902647606380533 -2.972614998298519 8.222053984136918 '
            '9.323339236176201">\n<path stroke-width="0.09" stroke="black" fill="none" d="M3.37 -0.30 L-2.59 -2.83 L5.13 6.13 '
            'L3.07 4.29 L2.44 6.35 L5.16 -2.97 L5.57 -2.66 L4.30 -1.24 L5.63 2.41 Z" />\n<path stroke-width="0.09" stroke="black" '
            'fill="none" d="M0.30 0.42 L0.03 0.12 L0.67 0.65 L0.62 0.38" />\n</svg>\n')
    assert triangulated.contour_sequences_to_svg(seqs) == want


def test_morph_triangle_orientation_agrees_with_reference_on_a_smooth_field():
    """MorphTriangles.orient_triangles (morph_geometry.py:40-67, time-compatible components, max-x rule at mid-life)
    on the reference's own morph triangles of the smooth 4D golden, windings scrambled: the reference's winding comes
    back for all but a handful of triangles (its DFS is order-dependent where small components tie)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "mp4d_morph7.npz"))
    tris = g["morph_triangles"]
    rng = np.random.default_rng(1)
    scrambled = np.array([t[::-1] if rng.random() < 0.5 else t for t in tris])
    mt = morph_geometry.MorphTriangles(g["morph_points"], g["morph_segments"], scrambled)
    mt.orient_triangles()

    def rot(t):
        t = tuple(int(i) for i in t)
        k = t.index(min(t))
        return t[k:] + t[:k]
    want, got = set(rot(t) for t in tris), set(rot(t) for t in mt.triangle_segment_indices)
    assert len(got) == len(want) == len(tris)
    assert len(want & got) >= 0.995 * len(tris)
    assert set(tuple(sorted(t)) for t in want) == set(tuple(sorted(t)) for t in got)


def _soup(field, z):
    from oracle import mt2d
    r = mt2d.extract_level(field, z)
    return r["seg_keys"], r["pos"][np.searchsorted(r["keys"], r["seg_keys"])]


def test_vectorised_chaining_equals_the_walk():
    """chain_segments ranks the contours by pointer jumping over darts (rank_darts); the result must be the per-vertex
    walk's (triangulated.py:236-293 with a fixed order): same polylines, same order, same start and direction."""
    from contourist_b200 import triangulated
    rng = np.random.default_rng(4)
    n_poly = 0
    for n in (9, 33, 120, 257):
        g = np.linspace(-2, 2, n)
        X, Y = np.meshgrid(g, g, indexing="ij")
        f = np.sqrt(np.sin(3 * X + Y * Y) ** 2 + np.cos(4 * Y + X * X) ** 2) + 0.08 * rng.standard_normal((n, n))
        for z in (0.25, 0.6, 1.0):
            k2, p2 = _soup(f, z)
            if len(k2) == 0:
                assert triangulated.chain_segments(k2, p2) == []
                continue
            uk, pts, e, deg = triangulated._graph(k2, p2)
            assert deg.max() <= 2
            want = triangulated._chain_walk(pts, e, deg)
            got = triangulated.chain_segments(k2, p2)
            assert len(got) == len(want)
            for (ca, pa), (cb, pb) in zip(got, want):
                assert ca == cb and np.array_equal(pa, pb)
            n_poly += len(got)
    assert n_poly > 500


def test_chaining_with_junctions_takes_the_walk():
    "a sample exactly on the level gives its keys more than two neighbours: the fixed-order walk decides"
    from contourist_b200 import triangulated
    rng = np.random.default_rng(1)
    f = rng.integers(-1, 2, size=(15, 15)).astype(float)
    k2, p2 = _soup(f, 0.0)
    uk, pts, e, deg = triangulated._graph(k2, p2)
    assert deg.max() > 2
    got = triangulated.chain_segments(k2, p2)
    want = triangulated._chain_walk(pts, e, deg)
    assert len(got) == len(want) and all(a == b and np.array_equal(p, q) for (a, p), (b, q) in zip(got, want))


def test_crossing_scan_with_skip_and_plateaus():
    """_segments._scan_samples restates grid_field.py:64-84 for any skip; against a literal transcription of the loop."""
    from contourist_b200 import _segments
    rng = np.random.default_rng(2)
    arr = rng.standard_normal((9, 8, 10))
    arr[2:5, 3:6, 4:8] = 0.3 + 1e-7 * rng.standard_normal((3, 3, 4))          # an np.allclose plateau on the level
    for skip in (1, 2, 3):
        fmax, fmin, p, q = _segments._scan_samples(arr, 0.3, skip)
        want = set()
        n = [s - 1 for s in arr.shape]
        for i in range(0, n[0], skip):
            for j in range(0, n[1], skip):
                for k in range(0, n[2], skip):
                    for index in range(1, 8):
                        o = [((index >> s) & 1) * skip for s in range(3)]
                        v1 = (i + o[0], j + o[1], k + o[2])
                        if all(v1[a] < arr.shape[a] for a in range(3)) and (arr[i, j, k] - 0.3) * (arr[v1] - 0.3) < 0:
                            want.add(((i, j, k), v1))
        got = set((tuple(int(x) for x in a), tuple(int(x) for x in b)) for a, b in zip(p, q))
        assert got == want and len(want) > 0
    from oracle import mt3d
    _, _, keys = mt3d.crossing_segments(arr, 0.3)
    _, _, p, q = _segments._scan_samples(arr, 0.3, 1)
    assert len(keys) == len(p)


def test_orient_triangles_honours_compatible_triangle_test():
    """surface_geometry.py:52-56: orient_triangles(compatible_triangle_test) -- the reference's first positional
    parameter -- vetoes links exactly like the vectorised link_filter."""
    from contourist_b200 import surface_geometry
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [2, 0, 0], [2, 1, 0]], dtype=float)
    T = np.array([[0, 1, 2], [1, 2, 3], [1, 4, 3], [4, 5, 3]])         # a strip of four triangles, mixed windings
    seen = []

    def veto_middle(t1, t2):
        seen.append((t1, t2))
        return not ({1, 3} <= set(t1) and {1, 3} <= set(t2))           # cut the strip at the edge (1, 3)
    a = surface_geometry.SurfaceGeometry(V, T).orient_triangles(veto_middle)
    mask = lambda k1, k2: np.array([not ({1, 3} <= set(T[i]) and {1, 3} <= set(T[j])) for i, j in zip(k1, k2)])
    b = surface_geometry.SurfaceGeometry(V, T).orient_triangles(link_filter=mask)
    c = surface_geometry.SurfaceGeometry(V, T).orient_triangles()
    assert a == b and len(seen) == 3 and all(isinstance(t, tuple) for pair in seen for t in pair)
    assert len(c) == len(a) == 4
