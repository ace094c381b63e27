"""GPU parity: the CUDA 4D pentatope + morph path (through the C ABI) against the numpy oracle and reference goldens."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mp4d

pytestmark = pytest.mark.gpu
FILES = sorted(glob.glob(os.path.join(GOLDEN, "mp4d_*.npz")))


def compare(engine, field, value, geom64=True, morph=True, origin=(0,) * 4, delta=(1,) * 4):
    from contourist_b200 import engine as E
    flags = E.WANT_KEYS | E.WANT_CODES | E.WANT_MINMAX | (E.GEOM_F64 if geom64 else 0) | (E.MORPH if morph else 0)
    c = engine.mp4d_run(field, value, origin=origin, delta=delta, flags=flags)
    o = engine.mp4d_fetch()
    gd = np.float64 if geom64 else np.float32
    r = mp4d.extract(field, value, gd)
    assert c.n_verts == len(r["keys"]) and c.n_tets == len(r["tets"]) and c.n_active_cells == len(r["cells"])
    assert c.fmin == field.min() and c.fmax == field.max()
    assert np.array_equal(o["keys"], r["keys"])
    assert np.array_equal(o["lowmin"], r["lowmin"])
    assert np.array_equal(np.sort(o["tets"], axis=1), np.sort(r["tets"], axis=1))      # same order, same split rule
    order = np.argsort(o["cells"])
    assert np.array_equal(o["cells"][order], r["cells"])
    assert np.array_equal(o["codes"][order], r["codes"])
    world = (r["pos"].astype(gd) * np.asarray(delta, gd) + np.asarray(origin, gd)).astype(gd)
    if geom64:
        assert np.array_equal(o["verts"], world)
    else:
        np.testing.assert_allclose(o["verts"], world, rtol=1e-4, atol=1e-5)
    if morph:
        r64 = r if geom64 else mp4d.extract(field, value, np.float64)
        corner = np.array(field.shape) - 1
        bp = mp4d.bin_times(r64["pos"], corner[3])
        assert np.array_equal(o["morph_verts"], bp)
        keep = mp4d.drop_instant(bp, r64["tets"]) & ~mp4d.tiny_mask(bp, r64["tets"], corner)
        assert np.array_equal(o["keep"].astype(bool), keep)
        segs, tris = mp4d.morph_triangles(bp, r64["tets"][keep])
        mt = o["morph_tris"].astype(np.int64)
        assert (bp[mt[:, :, 0], 3] <= bp[mt[:, :, 1], 3]).all()                        # low t first
        skey = np.minimum(mt[:, :, 0], mt[:, :, 1]) * (len(bp) + 1) + np.maximum(mt[:, :, 0], mt[:, :, 1])
        got = set(frozenset(int(x) for x in row) for row in skey)
        s2 = np.sort(segs, axis=1)
        okey = s2[:, 0] * (len(bp) + 1) + s2[:, 1]
        exp = set(frozenset(int(okey[x]) for x in t) for t in tris)
        assert got == exp
        assert c.t_min == bp[:, 3].min() and c.t_max == bp[:, 3].max()
    return c, o, r


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_golden_fields(engine, path, dtype):
    g = np.load(path)
    compare(engine, g["field"].astype(dtype), float(g["value"]), geom64=(dtype == np.float64))


@pytest.mark.parametrize("shape", [(2, 2, 2, 2), (3, 4, 5, 35), (5, 4, 3, 66), (6, 6, 6, 6)])
def test_random_ragged(engine, shape):
    rng = np.random.default_rng(sum(shape))
    compare(engine, rng.standard_normal(shape), 0.1, morph=(np.prod(shape) < 700))


def test_morph_field_and_world_transform(engine):
    n, nt = 13, 5
    g = np.linspace(-2, 2, n)
    t = np.linspace(0, 1, nt)
    X, Y, Z, T_ = np.meshgrid(g, g, g, t, indexing="ij")
    f = (T_ * 3 * np.sqrt(X * X + Z * Z) + (1 - T_) * 3 * np.sqrt((1 - np.sqrt(X * X + Y * Y)) ** 2 + Z * Z)).astype(np.float32)
    compare(engine, f, 1.2, geom64=True, morph=False, origin=(-2, -2, -2, 0), delta=(1 / 3.0, 1 / 3.0, 1 / 3.0, 0.25))


def test_equalities_and_plateau(engine):
    rng = np.random.default_rng(21)
    ints = rng.integers(-1, 2, size=(4, 5, 4, 34)).astype(np.float64)
    compare(engine, ints, 0.0, morph=False)
    shape = (4, 4, 4, 33)
    plate = np.where(rng.random(shape) < 0.7, 0.3 + 1e-7 * rng.standard_normal(shape), 0.3 + 0.4 * rng.standard_normal(shape))
    c, o, r = compare(engine, plate, 0.3, morph=False)
    assert ((r["codes"] & 32) != 0).any()


def test_empty(engine):
    c = engine.mp4d_run(np.zeros((3, 3, 3, 3)), 1.0)
    assert c.n_verts == 0 and c.n_tets == 0


def test_medium_size_properties(engine):
    """64^3 x 16 morph field: every tet has 4 distinct vertices, ids in range, every vertex used, and the
    tetrahedra form a closed 3-manifold away from the domain boundary (each interior triangular face is shared by
    exactly two tetrahedra)."""
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import synthetic
    f = synthetic.morph4d(48, 12, device="cuda")
    torch.cuda.synchronize()          # the engine has its own stream: the field must be complete before it reads it
    c = engine.mp4d_run(f.data_ptr(), 1.2, shape=tuple(f.shape), dtype=np.float32, flags=E.GEOM_F64)
    o = engine.mp4d_fetch()
    t = np.sort(o["tets"].astype(np.int64), axis=1)
    assert c.n_tets > 100000
    assert t.min() >= 0 and t.max() < c.n_verts
    assert (np.diff(t, axis=1) > 0).all()
    assert np.array_equal(np.unique(t), np.arange(c.n_verts))
    faces = np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]])
    code = (faces[:, 0] * c.n_verts + faces[:, 1]) * c.n_verts + faces[:, 2]
    u, cnt = np.unique(code, return_counts=True)
    assert cnt.max() <= 2
    # faces used once lie on the domain boundary: all 3 vertices share a boundary coordinate
    lone = u[cnt == 1]
    a = lone // (c.n_verts * c.n_verts)
    b = (lone // c.n_verts) % c.n_verts
    cc = lone % c.n_verts
    P = o["verts"]
    hi = np.array(f.shape) - 1
    on = np.zeros(len(lone), dtype=bool)
    for ax in range(4):
        for bound in (0.0, float(hi[ax])):
            on |= (P[a, ax] == bound) & (P[b, ax] == bound) & (P[cc, ax] == bound)
    assert on.all()
    del f
    torch.cuda.empty_cache()


SEEDED4D = sorted(glob.glob(os.path.join(GOLDEN, "seeded4d_*.npz")))


@pytest.mark.parametrize("path", SEEDED4D, ids=[os.path.basename(f) for f in SEEDED4D])
def test_facade_4d_seeded_equals_reference_tracker(engine, path):
    """GridContour4D with explicit seed segments returns what the reference's tracker reaches from them (pentatopes.py:92-106,
    tetrahedral.py:396-469 with the 80 offsets): same start hypervoxels, same reached hypervoxels, same edge keys, same
    number of tetrahedra as the run of the unmodified reference; tetrahedra and positions equal to the oracle's."""
    from contourist_b200 import pentatopes
    g = np.load(path)
    field, value = g["field"], float(g["value"])
    seeds = [[tuple(a), tuple(b)] for a, b in g["seeds"].tolist()]
    G = pentatopes.GridContour4D([s - 1 for s in field.shape], field, value, seeds)
    o = G.find_tetrahedra()
    assert [tuple(v) for v in G.start_voxels] == [tuple(v) for v in g["initial"].tolist()]
    assert np.array_equal(np.argwhere(G.selected_voxels), g["voxels"])
    r = mp4d.extract_seeded(field, value, g["seeds"])
    assert np.array_equal(o["keys"], r["keys"]) and len(o["tets"]) == int(g["n_tets"]) == len(r["tets"])
    assert np.array_equal(np.sort(o["tets"], axis=1), np.sort(r["tets"], axis=1))
    assert np.array_equal(o["verts"], r["pos"])
    full = pentatopes.GridContour4D([s - 1 for s in field.shape], field, value, None).find_tetrahedra()
    assert (len(o["tets"]) < len(full["tets"])) == ("both" not in path)
    # the morph triangles of the selection are the full scan's restricted to it
    mt = G.collect_morph_triangles()
    assert len(mt.triangle_segment_indices) > 0 and mt.points4d.shape == (len(o["keys"]), 4)
    if "both" in path:
        assert len(o["morph_tris"]) == len(full["morph_tris"])
