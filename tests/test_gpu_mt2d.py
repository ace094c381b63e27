"""GPU parity: the CUDA 2D multi-level path (through the C ABI) against the numpy oracle and reference goldens."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mt2d

pytestmark = pytest.mark.gpu
FILES = sorted(glob.glob(os.path.join(GOLDEN, "mt2d_*.npz")))


def compare(engine, field, levels, geom64=True, origin=(0, 0), delta=(1, 1)):
    from contourist_b200 import engine as E
    flags = (E.GEOM_F64 if geom64 else 0) | E.WANT_MINMAX
    c = engine.mt2d_run(field, levels, origin=origin, delta=delta, flags=flags)
    o = engine.mt2d_fetch()
    gd = np.float64 if geom64 else np.float32
    assert c.fmin == field.min() and c.fmax == field.max()
    total = 0
    for li, z in enumerate(levels):
        r = mt2d.extract_level(field, z, gd)
        sel = o["level"] == li
        sk = np.sort(o["keys"][sel], axis=1)
        # segments: bit-exact as a set (and no duplicates)
        uniq = np.unique(sk, axis=0)
        assert len(uniq) == len(sk)
        assert np.array_equal(uniq, r["seg_keys"])
        total += len(sk)
        # keys + positions
        k = o["keys"][sel].reshape(-1)
        p = o["pos"][sel].reshape(-1, 2)
        uk, first = np.unique(k, return_index=True)
        assert np.array_equal(uk, r["keys"])
        world = (r["pos"].astype(gd) * np.asarray(delta, gd) + np.asarray(origin, gd)).astype(gd)
        if geom64:
            assert np.array_equal(p[first], world)
        else:
            np.testing.assert_allclose(p[first], world, rtol=1e-4, atol=1e-5)
        # a key has the same position wherever it appears
        lut = dict(zip(uk.tolist(), range(len(uk))))
        idx = np.array([lut[x] for x in k.tolist()], dtype=np.int64)
        assert np.array_equal(p, p[first][idx])
    assert total == c.n_segments
    return c, o


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_golden_fields(engine, path, dtype):
    g = np.load(path)
    compare(engine, g["field"].astype(dtype), g["levels"], geom64=(dtype == np.float64))


@pytest.mark.parametrize("shape", [(2, 2), (3, 300), (300, 3), (65, 65), (40, 257), (33, 513)])
def test_random_ragged(engine, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    f = rng.standard_normal(shape)
    compare(engine, f, [-1.0, -0.2, 0.0, 0.3, 1.5])


def test_levels_equal_to_samples_and_world_transform(engine):
    rng = np.random.default_rng(2)
    f = rng.integers(0, 5, size=(37, 70)).astype(np.float32)
    compare(engine, f, [0.0, 1.0, 2.5, 3.0, 4.0], geom64=False, origin=(-2.0, 1.0), delta=(0.25, 0.5))
    compare(engine, f.astype(np.float64), [1.0, 2.0, 7.0], origin=(-2.0, 1.0), delta=(0.25, 0.5))


def test_sixteen_levels_oscillatory(engine):
    n = 129
    g = np.linspace(-2, 2, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    f = np.sqrt(np.sin(3 * X + Y * Y) ** 2 + np.cos(4 * Y + X * X) ** 2).astype(np.float32)
    levels = mt2d.linear_levels(f, 17)
    assert len(levels) == 16
    compare(engine, f, levels, geom64=True)


def test_empty_and_errors(engine):
    c = engine.mt2d_run(np.zeros((10, 10)), [1.0, 2.0])
    assert c.n_segments == 0
    with pytest.raises(ValueError):
        engine.mt2d_run(np.zeros((10, 10)), [2.0, 1.0])
    with pytest.raises(ValueError):
        engine.mt2d_run(np.zeros((1, 10)), [1.0])


def test_row_band_sharding(engine):
    from contourist_b200 import engine as E
    rng = np.random.default_rng(9)
    f = rng.standard_normal((90, 130)).astype(np.float32)
    levels = [-0.5, 0.1, 0.9]
    engine.mt2d_run(f, levels, flags=E.GEOM_F64)
    ref = engine.mt2d_fetch()
    parts = []
    bounds = [0, 31, 64, 90]
    for a, b in zip(bounds[:-1], bounds[1:]):
        hi = min(b + 1, f.shape[0])
        engine.mt2d_run(np.ascontiguousarray(f[a:hi]), levels, flags=E.GEOM_F64, i_lo=0, i_hi=b - a, row_offset=a)
        parts.append(engine.mt2d_fetch())
    for name in ("level", "keys", "pos"):
        assert np.array_equal(np.concatenate([p[name] for p in parts]), ref[name])


def test_full_size_properties_16384(engine):
    """BASELINE config 2: 16384^2 fp32, 16 levels.  Properties: every key is an end of exactly 2 segments
    except on the domain boundary (contours are closed or end on the boundary); positions agree per key."""
    import torch
    from contourist_b200 import synthetic
    n = 16384
    f = synthetic.field2d(n, device="cuda")
    torch.cuda.synchronize()          # the engine has its own stream: the field must be complete before it reads it
    mn, mx = float(f.min()), float(f.max())
    levels = [(mx - mn) / 17 * i for i in range(1, 17)]
    c = engine.mt2d_run(f.data_ptr(), levels, shape=(n, n), dtype=np.float32)
    o = engine.mt2d_fetch()
    assert c.n_segments > 1000000
    lvl = np.repeat(o["level"].astype(np.uint64), 2)
    k = o["keys"].reshape(-1)
    tag = (k << np.uint64(5)) | lvl               # key values use < 2^40 here
    u, cnt = np.unique(tag, return_counts=True)
    assert cnt.max() <= 2
    lone = u[cnt == 1] >> np.uint64(5)
    lin = (lone >> np.uint64(3)).astype(np.int64)
    i, j = lin // n, lin % n
    d = ((lone >> np.uint64(1)) & np.uint64(3)).astype(np.int64)
    on_edge = ((i == 0) & (d == 1)) | ((j == 0) & (d == 2)) | ((i == n - 1) & (d == 1)) | ((j == n - 1) & (d == 2))
    assert on_edge.all()
    del f
    torch.cuda.empty_cache()


def _host_chain(o, li):
    from contourist_b200 import triangulated
    sel = o["level"] == li
    return triangulated.chain_segments(o["keys"][sel], o["pos"][sel])


@pytest.mark.parametrize("geom64", [True, False])
def test_device_polylines_equal_host_chaining(engine, geom64):
    """ctr_mt2d_polylines (list ranking over darts on the device) against triangulated.chain_segments on the fetched
    segment soup (itself equal to the per-vertex walk, tests/test_facade_cpu.py): same polylines in the same order,
    same start, same direction, same points."""
    from contourist_b200 import engine as E
    rng = np.random.default_rng(8)
    total = 0
    for n0, n1 in ((9, 9), (65, 40), (200, 333)):
        g0, g1 = np.linspace(-2, 2, n0), np.linspace(-2, 2, n1)
        X, Y = np.meshgrid(g0, g1, indexing="ij")
        f = np.sqrt(np.sin(3 * X + Y * Y) ** 2 + np.cos(4 * Y + X * X) ** 2) + 0.07 * rng.standard_normal((n0, n1))
        if not geom64:
            f = f.astype(np.float32)
        levels = [0.2, 0.45, 0.7, 0.95, 1.3]
        engine.mt2d_run(f, levels, origin=(-1.0, 2.0), delta=(0.5, 0.25), flags=E.GEOM_F64 if geom64 else 0)
        polys = engine.mt2d_polylines()
        o = engine.mt2d_fetch()
        assert polys is not None and len(polys) == len(levels)
        for li in range(len(levels)):
            want = _host_chain(o, li)
            got = polys[li]
            assert len(got) == len(want)
            for (ca, pa), (cb, pb) in zip(got, want):
                assert ca == cb and np.array_equal(pa, pb)
            total += len(got)
    assert total > 300


def test_device_polylines_report_junctions(engine):
    rng = np.random.default_rng(1)
    f = rng.integers(-1, 2, size=(15, 15)).astype(float)
    engine.mt2d_run(f, [0.0])
    assert engine.mt2d_polylines() is None                    # a sample on the level: the host walk decides
    engine.mt2d_run(f, [0.5])
    polys = engine.mt2d_polylines()
    o = engine.mt2d_fetch()
    want = _host_chain(o, 0)
    assert len(polys[0]) == len(want) and all(a == b and np.array_equal(p, q) for (a, p), (b, q) in zip(polys[0], want))


def test_more_than_64_levels_through_the_facade(engine):
    from contourist_b200 import multiple_2d_contour

    def f(x, y):
        return np.sqrt(np.sin(3 * x + y * y) ** 2 + np.cos(4 * y + x * x) ** 2)
    C = multiple_2d_contour.Linear2DContour(-2, -2, 2, 2, 0.05, 0.05, f, breakpoints=100)
    assert len(C.values) == 99
    D = C.get_contours_dictionary()
    assert sorted(D) == sorted(C.values)
    arr = C.grid.samples(0)
    for value in (C.values[3], C.values[70], C.values[98]):
        r = mt2d.extract_level(arr, value)
        n_keys = len(np.unique(np.concatenate([np.round(p, 12) for _, p in D[value]]), axis=0)) if D[value] else 0
        assert (n_keys > 0) == (len(r["keys"]) > 0)
        if D[value]:
            allp = set(map(tuple, np.concatenate([p for _, p in D[value]])))
            refp = set(map(tuple, r["pos"] * 0.05 - 2.0))
            assert allp <= refp and len(refp) - len(allp) <= 0.01 * len(refp)       # consecutive np.allclose points are dropped
