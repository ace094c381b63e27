"""GPU parity: the CUDA 3D path (through the C ABI) against the numpy oracle and the reference goldens."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mt3d

pytestmark = pytest.mark.gpu

FILES = sorted(glob.glob(os.path.join(GOLDEN, "mt3d_*.npz")))


def run_and_compare(engine, field, value, geom64=True, origin=(0, 0, 0), delta=(1, 1, 1)):
    from contourist_b200 import engine as E
    flags = E.WANT_KEYS | E.WANT_CODES | E.WANT_NORMALS | E.WANT_MINMAX | (E.GEOM_F64 if geom64 else 0)
    c = engine.mt3d_run(field, value, origin=origin, delta=delta, flags=flags)
    raw = engine.mt3d_fetch()
    assert len(np.unique(raw["keys"])) == c.n_verts                      # one vertex per edge key (dedup is exact)
    o = E.canonical_mesh(raw)                                            # canonical sort: vertices by edge key
    gd = np.float64 if geom64 else np.float32
    r = mt3d.extract(field, value, gd)
    assert c.n_verts == len(r["keys"]) and c.n_tris == len(r["tris"])
    # bit-exact: active set, case codes, edge keys, key orientation, topology
    assert np.array_equal(o["keys"], r["keys"])
    assert np.array_equal(o["lowmin"], r["lowmin"])
    assert np.array_equal(np.sort(o["tris"], axis=1), np.sort(r["tris"], axis=1))
    order = np.argsort(o["cells"])
    assert np.array_equal(o["cells"][order], r["cells"])
    assert np.array_equal(o["codes"][order], r["codes"])
    assert c.n_active_cells == len(r["cells"])
    mx, mn, ncross = mt3d.crossing_segments(field, value, count_only=True)
    assert c.n_crossings == ncross and c.fmin == mn and c.fmax == mx
    # geometry
    world = (r["pos"].astype(gd) * np.asarray(delta, gd) + np.asarray(origin, gd)).astype(gd)
    if geom64:
        assert np.array_equal(o["verts"], world)          # fp64 mode: 0 ulp (north_star asks 1e-6 rel)
    else:
        np.testing.assert_allclose(o["verts"], world, rtol=1e-4, atol=1e-5)   # fp32 mode: 1e-4 rel
    nr = mt3d.normals(field, value, r["keys"], gd, delta=delta)
    np.testing.assert_allclose(o["normals"], nr, rtol=1e-9 if geom64 else 1e-4, atol=1e-11 if geom64 else 1e-4)
    return c, o, r


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_golden_fields(engine, path, dtype):
    g = np.load(path)
    field = g["field"].astype(dtype)
    c, o, r = run_and_compare(engine, field, float(g["value"]), geom64=(dtype == np.float64))
    if dtype == np.float64:
        # straight against the reference's own output (not via the oracle)
        n0, n1, n2 = field.shape
        klow, khigh = g["key_low"], g["key_high"]
        pmin = np.minimum(klow, khigh)
        d = np.maximum(klow, khigh) - pmin
        gk = ((((pmin[:, 0] * n1 + pmin[:, 1]) * n2 + pmin[:, 2]).astype(np.uint64) << np.uint64(3))
              | (d[:, 0] * 4 + d[:, 1] * 2 + d[:, 2]).astype(np.uint64))
        srt = np.argsort(gk)
        assert np.array_equal(gk[srt], o["keys"])
        assert np.array_equal(g["key_pos"][srt], o["verts"])
        assert len(g["tris"]) == c.n_tris


def test_c1_sphere65_against_full_reference_run(engine):
    """BASELINE configs[0] through the C ABI, straight against the reference's own run (tests/golden/c1_sphere65.npz)."""
    from conftest import c1_golden
    g = c1_golden()
    c, o, r = run_and_compare(engine, g["field"], 0.5)
    # 9,656 surface voxels in the reference; 24 of them only touch the level set in a corner with f == value exactly
    # (border_voxel is inclusive, tetrahedral.py:383-394) and emit nothing
    assert (c.n_active_cells, c.n_verts, c.n_tris) == (9632, 28778, 57552)
    n1 = n2 = 66
    pmin = np.minimum(g["key_low"], g["key_high"])
    d = np.maximum(g["key_low"], g["key_high"]) - pmin
    gk = ((((pmin[:, 0] * n1 + pmin[:, 1]) * n2 + pmin[:, 2]).astype(np.uint64) << np.uint64(3))
          | (d[:, 0] * 4 + d[:, 1] * 2 + d[:, 2]).astype(np.uint64))
    srt = np.argsort(gk)
    assert np.array_equal(gk[srt], o["keys"])
    assert np.array_equal(g["key_pos"][srt], o["verts"])                  # 0 ulp in fp64 mode
    gv = g["voxels"]
    assert len(gv) == 9656 and np.isin(o["cells"], (gv[:, 0] * 65 + gv[:, 1]) * 65 + gv[:, 2]).all()


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 2, 33), (33, 17, 70), (5, 64, 31), (40, 40, 40), (9, 9, 129)])
def test_random_fields_ragged_shapes(engine, shape):
    rng = np.random.default_rng(sum(shape))
    run_and_compare(engine, rng.standard_normal(shape), 0.05)


def test_world_transform_and_f32_field_f64_geometry(engine):
    g = np.linspace(-1, 1, 49)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    f = (X * X + 0.5 * Y * Y + Z * Z + 0.1 * np.sin(7 * X)).astype(np.float32)
    run_and_compare(engine, f, 0.5, geom64=True, origin=(-1.0, -1.0, -1.0), delta=(1 / 24.0, 1 / 24.0, 1 / 24.0))
    run_and_compare(engine, f, 0.5, geom64=False, origin=(-1.0, 2.0, 0.5), delta=(0.25, 0.5, 2.0))


def test_exact_equality_and_plateaus(engine):
    rng = np.random.default_rng(3)
    ints = rng.integers(-2, 3, size=(21, 19, 37)).astype(np.float32)       # many samples == isovalue
    run_and_compare(engine, ints, 0.0, geom64=False)
    run_and_compare(engine, ints.astype(np.float64), 1.0)
    shape = (12, 12, 40)                                                    # patchy allclose(value) plateau
    plate = np.where(rng.random(shape) < 0.6, 0.3 + 1e-7 * rng.standard_normal(shape),
                     0.3 + 0.4 * rng.standard_normal(shape))
    c, o, r = run_and_compare(engine, plate, 0.3)
    skipped = ((r["codes"][:, None] >> (5 * np.arange(6, dtype=np.uint32))[None, :]) & 16).any()
    assert skipped


def test_empty_and_constant_fields(engine):
    from contourist_b200 import engine as E
    for f in (np.zeros((8, 8, 8)), np.full((4, 5, 6), 2.0, dtype=np.float32)):
        c = engine.mt3d_run(f, 1.0, flags=E.WANT_KEYS)
        assert c.n_verts == 0 and c.n_tris == 0 and c.n_active_cells == 0
        o = engine.mt3d_fetch()
        assert o["verts"].shape == (0, 3) and o["tris"].shape == (0, 3)
    c = engine.mt3d_run(np.full((4, 4, 4), 1.0), 1.0)     # f == v everywhere: all "high", all allclose
    assert c.n_tris == 0


def test_triangles_wound_toward_high_side(engine):
    from contourist_b200 import engine as E
    g = np.linspace(-1, 1, 41)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    f = X * X + Y * Y + Z * Z                      # increases outward: normals must point outward
    engine.mt3d_run(f, 0.5, flags=E.GEOM_F64 | E.WANT_NORMALS)
    o = engine.mt3d_fetch()
    P, t = o["verts"], o["tris"]
    fn = np.cross(P[t[:, 1]] - P[t[:, 0]], P[t[:, 2]] - P[t[:, 0]])
    cen = (P[t[:, 0]] + P[t[:, 1]] + P[t[:, 2]]) / 3 - 20.0
    area = np.linalg.norm(fn, axis=1)
    big = area > 1e-9                      # samples exactly on the isovalue give zero-area triangles
    assert big.sum() > 0.9 * len(t)
    assert ((fn * cen).sum(axis=1)[big] > 0).all()
    gn = o["normals"][t[:, 0]]
    assert ((fn * gn).sum(axis=1)[big] > 0).all()


def test_bad_arguments_raise(engine):
    with pytest.raises(ValueError):
        engine.mt3d_run(np.zeros((1, 4, 4)), 0.0)
    with pytest.raises(ValueError):
        engine.mt3d_run(np.zeros((4, 4, 4)), float("nan"))
    with pytest.raises(ValueError):
        engine.mt3d_run(np.zeros((4, 4, 4)), 0.0, delta=(1, 0, 1))


def test_slab_sharding_matches_single_run(engine):
    """Multi-GPU decomposition emulated on one device: N z-slabs with halo planes reproduce the single-run
    mesh exactly (global keys, positions, and triangles after adding each shard's vertex offset)."""
    from contourist_b200 import engine as E
    rng = np.random.default_rng(11)
    g = np.linspace(-1, 1, 50)
    X, Y, Z = np.meshgrid(g, g[:31], g[:37], indexing="ij")
    f = (np.sin(4 * X) * np.cos(3 * Y) + Z * Z + 0.05 * rng.standard_normal(X.shape)).astype(np.float32)
    flags = E.WANT_KEYS | E.WANT_NORMALS | E.GEOM_F64
    c0 = engine.mt3d_run(f, 0.2, flags=flags)
    ref = engine.mt3d_fetch()
    n0 = f.shape[0]
    for nshard in (2, 3, 7):
        bounds = [round(r * n0 / nshard) for r in range(nshard + 1)]
        keys, verts, normals, tris, voff = [], [], [], [], 0
        for r in range(nshard):
            a, b = bounds[r], bounds[r + 1]
            lo = max(a - 1, 0)
            hi = min(b + 2, n0)
            sub = np.ascontiguousarray(f[lo:hi])
            c = engine.mt3d_run(sub, 0.2, flags=flags, i_lo=a - lo, i_hi=b - lo, plane_offset=lo)
            o = engine.mt3d_fetch()
            assert o["tris"].min() >= 0
            keys.append(o["keys"]); verts.append(o["verts"]); normals.append(o["normals"])
            tris.append(o["tris"].astype(np.int64) + voff)
            voff += c.n_verts
        assert np.array_equal(np.concatenate(keys), ref["keys"])
        assert np.array_equal(np.concatenate(verts), ref["verts"])
        np.testing.assert_allclose(np.concatenate(normals), ref["normals"], rtol=1e-12, atol=1e-14)
        assert np.array_equal(np.concatenate(tris), ref["tris"].astype(np.int64))


def test_enqueue_finish_equals_run(engine):
    """ctr_mt3d_enqueue + ctr_mt3d_finish is ctr_mt3d_run split at the wait (also when the pools have to grow)."""
    from contourist_b200 import engine as E
    rng = np.random.default_rng(21)
    flags = E.WANT_KEYS | E.WANT_NORMALS | E.GEOM_F64
    for shape in ((20, 33, 40), (70, 64, 96)):           # the second is larger than anything this context has seen
        f = rng.standard_normal(shape)
        c0 = engine.mt3d_run(f, 0.1, flags=flags)
        ref = engine.mt3d_fetch()
        engine.mt3d_enqueue(f, 0.1, flags=flags)
        c1 = engine.mt3d_finish()
        out = engine.mt3d_fetch()
        assert (c1.n_verts, c1.n_tris, c1.n_active_cells, c1.n_crossings) == (c0.n_verts, c0.n_tris, c0.n_active_cells, c0.n_crossings)
        for name in ("keys", "lowmin", "verts", "normals", "tris"):
            assert np.array_equal(out[name], ref[name]), name
    big = rng.standard_normal((90, 80, 128))
    fresh = E.Engine(0)                                  # first call of a context: pools are guesses, finish() must redo
    fresh.mt3d_enqueue(big, 0.0, flags=flags)
    c2 = fresh.mt3d_finish()
    o2 = E.canonical_mesh(fresh.mt3d_fetch())
    c3 = engine.mt3d_run(big, 0.0, flags=flags)
    o3 = E.canonical_mesh(engine.mt3d_fetch())
    assert c2.n_verts == c3.n_verts and c2.n_tris == c3.n_tris
    for name in ("keys", "verts", "tris"):
        assert np.array_equal(o2[name], o3[name]), name
    with pytest.raises(E.EngineError):
        fresh.mt3d_finish()                              # nothing pending
    fresh.close()


def test_pools_grow_from_a_tiny_first_run(engine):
    """A context whose pools were sized by a tiny run meets volumes whose lists overflow them several times over: the
    speculative stages must stay inside the capacities (never read list entries behind them) and the redo must give the
    mesh of a context that has always been large."""
    from contourist_b200 import engine as E
    rng = np.random.default_rng(77)
    flags = E.WANT_KEYS | E.WANT_NORMALS
    small = E.Engine(0)
    small.mt3d_run(rng.standard_normal((6, 7, 9)), 0.0, flags=flags)
    for shape in ((40, 33, 70), (64, 96, 160), (150, 130, 257)):
        f = rng.standard_normal(shape).astype(np.float32)
        c0 = small.mt3d_run(f, 0.0, flags=flags)
        o0 = small.mt3d_fetch()
        c1 = engine.mt3d_run(f, 0.0, flags=flags)
        o1 = engine.mt3d_fetch()
        assert (c0.n_verts, c0.n_tris, c0.n_active_cells, c0.n_crossings) == (c1.n_verts, c1.n_tris, c1.n_active_cells, c1.n_crossings)
        for name in ("keys", "lowmin", "verts", "normals", "tris"):
            assert np.array_equal(o0[name], o1[name]), name
    small.close()


def test_alternating_fields_on_one_context(engine):
    """Two different fields of one shape run in turn on the same context: every buffer of a run is rewritten by the next
    at the same addresses.  The kernels of a run are chained with programmatic dependent launch and some start under the
    tail of the one in front; a kernel that read anything stale (a bit plane, a record, a list entry of the run before)
    would give another mesh than a context that has only ever seen that field."""
    from contourist_b200 import engine as E
    rng = np.random.default_rng(99)
    shape = (96, 80, 160)
    fa = rng.standard_normal(shape).astype(np.float32)
    g = np.linspace(-1, 1, 160)
    fb = (np.sin(7 * g)[None, None, :] * np.cos(5 * g[:80])[None, :, None] + g[:96, None, None] ** 2
          + 0.2 * rng.standard_normal(shape)).astype(np.float32)
    flags = E.WANT_KEYS | E.WANT_NORMALS
    refs = []
    for f in (fa, fb):
        fresh = E.Engine(0)
        c = fresh.mt3d_run(f, 0.1, flags=flags)
        refs.append((c.n_verts, c.n_tris, fresh.mt3d_fetch()))
        fresh.close()
    for rep in range(8):
        for f, (nv, nt, ref) in zip((fa, fb), refs):
            c = engine.mt3d_run(f, 0.1, flags=flags)
            out = engine.mt3d_fetch()
            assert (c.n_verts, c.n_tris) == (nv, nt)
            for name in ("keys", "lowmin", "verts", "normals", "tris"):
                assert np.array_equal(out[name], ref[name]), (name, rep)


def test_pipelined_host_extraction_equals_single_run(engine):
    """engine.mt3d_extract_host (slab-by-slab upload / extract / download on two contexts, global triangle ids via
    vert_id_base) returns exactly the mesh of one ctr_mt3d_run + ctr_mt3d_fetch."""
    from contourist_b200 import engine as E
    rng = np.random.default_rng(5)
    g = np.linspace(-1, 1, 61)
    X, Y, Z = np.meshgrid(g, g[:40], g[:70], indexing="ij")
    f = (np.sin(5 * X) * np.cos(3 * Y) + Z * Z + 0.05 * rng.standard_normal(X.shape)).astype(np.float32)
    flags = E.WANT_KEYS | E.WANT_NORMALS
    c = engine.mt3d_run(f, 0.2, origin=(1, 2, 3), delta=(0.5, 0.25, 2), flags=flags)
    ref = engine.mt3d_fetch()
    for nslabs in (1, 3, 4):
        for rep in range(2):                             # first call sizes its pools, the second runs overlapped
            tot, out = engine.mt3d_extract_host(f, 0.2, origin=(1, 2, 3), delta=(0.5, 0.25, 2), flags=flags, nslabs=nslabs)
            assert tot["n_verts"] == c.n_verts and tot["n_tris"] == c.n_tris
            assert tot["n_active_cells"] == c.n_active_cells and tot["n_crossings"] == c.n_crossings
            for name in ("keys", "lowmin", "verts", "normals", "tris"):
                assert np.array_equal(out[name], ref[name]), (name, nslabs, rep)


def test_pipelined_host_extraction_with_empty_slabs_on_a_new_context(engine):
    """The surface lies in the last planes only: the first slabs of a context's first call produce nothing (no output
    buffers exist yet), later calls run overlapped; the halo planes are uploaded once."""
    from contourist_b200 import engine as E
    f = np.full((48, 20, 40), -1.0, dtype=np.float32)
    f[40:] = np.linspace(-1, 1, 8 * 20 * 40, dtype=np.float32).reshape(8, 20, 40)
    flags = E.WANT_KEYS | E.WANT_NORMALS
    c = engine.mt3d_run(f, 0.25, flags=flags)
    ref = engine.mt3d_fetch()
    assert c.n_tris > 0
    fresh = E.Engine(0)
    for rep in range(2):
        tot, out = fresh.mt3d_extract_host(f, 0.25, flags=flags, nslabs=12)
        assert tot["slabs"][0] == (0, 0) and tot["n_verts"] == c.n_verts and tot["n_tris"] == c.n_tris
        for name in ("keys", "lowmin", "verts", "normals", "tris"):
            assert np.array_equal(out[name], ref[name]), (name, rep)
    fresh.close()


def test_full_size_properties_512(engine):
    """BASELINE config 3 size: 512^3 fp32 CT-like volume.  Size-independent properties: the mesh is a closed
    2-manifold away from the domain boundary (every interior edge shared by exactly 2 triangles with opposite
    direction), Euler-consistent counts, indices in range, and per-stage counts agree with a z-slab re-run."""
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import synthetic
    n = 512
    f = synthetic.ct_like(n, device="cuda")
    torch.cuda.synchronize()          # the engine has its own stream: the field must be complete before it reads it
    c = engine.mt3d_run(f.data_ptr(), 0.5, shape=(n, n, n), dtype=np.float32, flags=E.WANT_NORMALS | E.WANT_KEYS)
    o = engine.mt3d_fetch()
    t = o["tris"].astype(np.int64)
    assert t.min() >= 0 and t.max() < c.n_verts and c.n_tris > 1000000
    assert np.array_equal(np.unique(t), np.arange(c.n_verts))            # every vertex used
    assert len(np.unique(o["keys"])) == c.n_verts                        # one vertex per edge key
    own = (o["keys"] >> np.uint64(3)).astype(np.int64) // 32             # owner word of each vertex
    assert (np.diff(own) >= 0).all()                                     # ids are word-major
    # directed edges: each appears once; interior edges have their reverse
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    code = e[:, 0] * c.n_verts + e[:, 1]
    assert len(np.unique(code)) == len(code)
    rev = e[:, 1] * c.n_verts + e[:, 0]
    has_rev = np.isin(code, rev)
    # edges without a reverse lie on the domain boundary
    P = o["verts"]
    lone = e[~has_rev]
    pa, pb = P[lone[:, 0]], P[lone[:, 1]]
    on_face = ((pa == 0) & (pb == 0)) | ((pa == n - 1) & (pb == n - 1))
    assert on_face.any(axis=1).all()
    nn = np.linalg.norm(o["normals"], axis=1)
    assert np.all((np.abs(nn - 1) < 1e-3) | (nn == 0))
    # slab re-run (2 shards) reproduces counts
    half = n // 2
    c_a = engine.mt3d_run(f.data_ptr(), 0.5, shape=(half + 2, n, n), dtype=np.float32, flags=E.NO_GEOMETRY,
                          i_lo=0, i_hi=half)
    c_b = engine.mt3d_run(f[half - 1:].data_ptr(), 0.5, shape=(n - half + 1, n, n), dtype=np.float32,
                          flags=E.NO_GEOMETRY, i_lo=1, i_hi=n - half + 1, plane_offset=half - 1)
    assert c_a.n_verts + c_b.n_verts == c.n_verts
    assert c_a.n_tris + c_b.n_tris == c.n_tris
    assert c_a.n_active_cells + c_b.n_active_cells == c.n_active_cells
    del f
    torch.cuda.empty_cache()


def test_contexts_share_no_process_state(engine):
    """Kernel attributes (dynamic shared memory opt-in of the TMA bitplane kernel, of the 4D slicer) are per device and
    per kernel instance: a second context, a 4D run after a 3D run in the same context, and -- with two GPUs -- a
    context on the other device must all launch.  Shapes are TMA-eligible (rows a multiple of 32 samples)."""
    import torch
    from contourist_b200 import engine as E
    rng = np.random.default_rng(9)
    f3 = rng.standard_normal((6, 5, 64)).astype(np.float32)
    f4 = rng.standard_normal((4, 3, 3, 32)).astype(np.float32)
    want3 = mt3d.extract(f3, 0.1, np.float32)
    devices = [0] + ([1] if torch.cuda.device_count() > 1 else [])
    for dev in devices:
        a, b = E.Engine(dev), E.Engine(dev)
        try:
            for eng in (a, b, a):
                c = eng.mt3d_run(f3, 0.1, flags=E.WANT_KEYS)
                assert c.n_verts == len(want3["keys"]) and c.n_tris == len(want3["tris"])
                c4 = eng.mp4d_run(f4, 0.1, flags=E.WANT_KEYS | E.MORPH)
                assert c4.n_verts > 0
                c = eng.mt3d_run(f3, 0.1, flags=E.WANT_KEYS | E.WANT_MINMAX)
                assert np.array_equal(np.sort(eng.mt3d_fetch()["keys"]), want3["keys"])
        finally:
            a.close()
            b.close()
