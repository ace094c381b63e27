"""GPU parity ON THE CONFIGS THE NUMBERS ARE QUOTED ON (BASELINE.json configs[1..4]): the engine runs the exact bench
fields at full size and its output is compared with the numpy oracle on random sub-blocks (the oracle needs a second
per 56^3 block; the full fields would take it minutes).  A sub-block is cut out of the field with its own samples
only; what the oracle extracts from it is exactly what the engine extracted for the voxels / squares inside the block,
so the engine's output is filtered by location and must then be EQUAL: keys, key orientation, triangles / segments /
tetrahedra bit-exact, positions and normals within the fp32-mode tolerance of north_star (1e-4 relative).
"""
import numpy as np
import pytest

from oracle import mp4d, mt2d, mt3d

pytestmark = pytest.mark.gpu

B3 = 56          # samples per axis of a 3D sub-block
NBLOCKS = 8


def _blocks(rng, shape, size, n, lo=None, hi=None):
    """n random block origins inside `shape` (the first two pinned to the low and the high corner of the volume)."""
    lo = [0] * len(shape) if lo is None else lo
    hi = [s - b for s, b in zip(shape, size)] if hi is None else hi
    out = [tuple(lo), tuple(hi)]
    while len(out) < n:
        out.append(tuple(int(rng.integers(l, h + 1)) for l, h in zip(lo, hi)))
    return out


def _decode3(keys, n1, n2):
    lin = (keys >> np.uint64(3)).astype(np.int64)
    d = (keys & np.uint64(7)).astype(np.int64)
    p = np.stack([lin // (n1 * n2), (lin // n2) % n1, lin % n2], axis=1)
    return p, p + np.stack([(d >> 2) & 1, (d >> 1) & 1, d & 1], axis=1)


def _check_blocks_3d(field_t, iso, out, n_verts, shape_global, plane0, blocks, size, voxel_plane_hi=None):
    """field_t: torch tensor holding planes [plane0, plane0 + field_t.shape[0]) of the global volume."""
    n0, n1, n2 = shape_global
    keys, tris = out["keys"], out["tris"].astype(np.int64)
    order = np.argsort(keys)
    skeys = keys[order]
    # triangles whose three vertices this run owns (a sharded run references the next shard's first plane by id)
    local = (tris < n_verts).all(axis=1)
    tk = keys[tris[local]]
    pv = _decode3(keys, n1, n2)[0].astype(np.int32)                                 # per vertex: min endpoint of its edge
    vox = pv[tris[local]].min(axis=1)                 # [T, 3]: voxel origin = corner A, an endpoint in every triangle
    n_tri_checked = n_pts_checked = 0
    for org in blocks:
        org = np.array(org)
        sub = field_t[org[0] - plane0:org[0] - plane0 + size[0], org[1]:org[1] + size[1], org[2]:org[2] + size[2]].cpu().numpy()
        assert sub.shape == tuple(size)
        r = mt3d.extract(sub, iso, np.float32)
        # local keys -> global keys
        pl, _ = _decode3(r["keys"], size[1], size[2])
        gl = (((pl[:, 0] + org[0]) * n1 + (pl[:, 1] + org[1])) * n2 + (pl[:, 2] + org[2])).astype(np.uint64)
        gkeys = (gl << np.uint64(3)) | (r["keys"] & np.uint64(7))
        want_tris = np.sort(gkeys[r["tris"]], axis=1)
        top = org + np.array(size) - 2                                              # last voxel origin inside the block
        sel = np.all((vox >= org) & (vox <= top), axis=1)
        if voxel_plane_hi is not None:
            sel &= vox[:, 0] <= voxel_plane_hi
        got_tris = np.sort(tk[sel], axis=1)
        a = np.unique(got_tris, axis=0)
        b = np.unique(want_tris, axis=0)
        assert len(a) == len(got_tris) and len(b) == len(want_tris)                 # no duplicate triangles
        assert np.array_equal(a, b)                                                 # topology: bit-exact
        # vertices of the block: present, same orientation, positions / normals in tolerance
        pos = np.searchsorted(skeys, gkeys)
        assert np.all(pos < len(skeys)) and np.array_equal(skeys[pos], gkeys)
        idx = order[pos]
        assert np.array_equal(out["lowmin"][idx], r["lowmin"])
        np.testing.assert_allclose(out["verts"][idx], r["pos"] + org.astype(np.float32), rtol=1e-4, atol=1e-4)
        # gradient normals use a 1-sample stencil: compare where both endpoints are interior to the block
        p0, p1 = _decode3(r["keys"], size[1], size[2])
        inner = np.all((p0 >= 1) & (p1 <= np.array(size) - 2), axis=1)
        nr = mt3d.normals(sub, iso, r["keys"], np.float32)
        np.testing.assert_allclose(out["normals"][idx][inner], nr[inner], rtol=1e-4, atol=2e-4)
        n_tri_checked += len(want_tris)
        n_pts_checked += len(gkeys)
    return n_tri_checked, n_pts_checked


def test_c3_ct512_subblocks_equal_oracle(engine):
    """BASELINE configs[2], the headline: synthetic.ct_like(512), isovalue 0.5, fp32 mesh + normals, exactly as bench.py."""
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import synthetic
    n = 512
    f = synthetic.ct_like(n)
    torch.cuda.synchronize()
    c = engine.mt3d_run(f.data_ptr(), 0.5, flags=E.WANT_NORMALS | E.WANT_KEYS, shape=(n, n, n), dtype=np.float32)
    out = engine.mt3d_fetch()
    assert (c.n_verts, c.n_tris, c.n_active_cells) == (3937166, 7874356, 1300935)    # the counts bench.py reports
    rng = np.random.default_rng(512)
    blocks = _blocks(rng, (n, n, n), (B3,) * 3, NBLOCKS)
    nt, nv = _check_blocks_3d(f, 0.5, out, c.n_verts, (n, n, n), 0, blocks, (B3,) * 3)
    assert nt > 20000 and nv > 10000                                                 # the blocks did contain surface


def test_c5_turbulence2048_slab_subblocks_equal_oracle(engine):
    """BASELINE configs[4]: a 48-owner-plane slab of the 2048^3 turbulence field (isovalue 0), run the way
    tools/bench_c5.py runs every rank's chunks (halo planes, i_lo / i_hi / plane_offset), global keys."""
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import sharding, synthetic
    n = 2048
    a, b = 1000, 1048
    lo, hi, kw = sharding.slab_with_halo(a, b, n)
    f = synthetic.turbulence(n, lo, hi, n_total=n)
    torch.cuda.synchronize()
    c = engine.mt3d_run(f.data_ptr(), 0.0, flags=E.WANT_NORMALS | E.WANT_KEYS, shape=(hi - lo, n, n), dtype=np.float32, **kw)
    out = engine.mt3d_fetch()
    assert c.n_verts > 5e6 and len(out["keys"]) == c.n_verts
    rng = np.random.default_rng(2048)
    size = (40, B3, B3)
    # blocks hold sample planes [a, b): voxel planes up to b-2 (the vertices of plane b belong to the next slab)
    blocks = _blocks(rng, (n, n, n), size, NBLOCKS, lo=[a, 0, 0], hi=[b - size[0], n - B3, n - B3])
    nt, nv = _check_blocks_3d(f, 0.0, out, c.n_verts, (n, n, n), lo, blocks, size)
    assert nt > 20000 and nv > 10000


def test_c4_morph128x64_subblocks_equal_oracle(engine):
    """BASELINE configs[3]: synthetic.morph4d(128, 64), isovalue 1.2, exactly as tools/bench_paths.py (+ keys)."""
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import synthetic
    n, nt = 128, 64
    f = synthetic.morph4d(n, nt)
    torch.cuda.synchronize()
    shape = (n, n, n, nt)
    c = engine.mp4d_run(f.data_ptr(), 1.2, shape=shape, dtype=np.float32, flags=E.WANT_KEYS)
    out = engine.mp4d_fetch()
    keys, tets = out["keys"], out["tets"].astype(np.int64)
    assert len(keys) == c.n_verts and len(tets) == c.n_tets and c.n_tets > 1e7
    assert np.all(keys[1:] > keys[:-1])                                              # 4D vertex ids are key ranks

    def decode(k, n1, n2, n3):
        lin = (k >> np.uint64(4)).astype(np.int64)
        d = (k & np.uint64(15)).astype(np.int64)
        p = np.stack([lin // (n1 * n2 * n3), (lin // (n2 * n3)) % n1, (lin // n3) % n2, lin % n3], axis=1)
        return p, d
    rng = np.random.default_rng(4)
    size = np.array((14, 14, 14, 33))
    pv, dv = decode(keys, n, n, nt)                                                   # per vertex: min endpoint, direction
    pvmax = (pv + np.stack([(dv >> 3) & 1, (dv >> 2) & 1, (dv >> 1) & 1, dv & 1], axis=1)).astype(np.int16)
    pv = pv.astype(np.int16)
    # the surface f = 1.2 lives around the torus / bar: block origins near vertices picked at random
    blocks = [np.clip(pv[q].astype(np.int64) - size // 2, 0, np.array(shape) - size) for q in rng.integers(0, len(keys), NBLOCKS)]
    checked = 0
    t0 = tets[:, 0]
    for org in blocks:
        sl = tuple(slice(int(o), int(o) + int(s_)) for o, s_ in zip(org, size))
        sub = f[sl].cpu().numpy()
        r = mp4d.extract(sub, 1.2, np.float32)
        pl, d = decode(r["keys"], int(size[1]), int(size[2]), int(size[3]))
        g = pl + org
        gkeys = ((((g[:, 0] * n + g[:, 1]) * n + g[:, 2]) * nt + g[:, 3]).astype(np.uint64) << np.uint64(4)) | d.astype(np.uint64)
        want = np.unique(np.sort(gkeys[r["tets"]], axis=1), axis=0)
        assert len(want) == len(r["tets"])
        # Engine tetrahedra with every vertex on an edge inside the block.  The four edges of a tetrahedron touch all
        # five corners of its pentatope, which include the hypervoxel's origin and far corner: such a tetrahedron
        # belongs to a hypervoxel inside the block, i.e. to the oracle's set -- and the other way round.
        top = org + size - 1
        cand = np.nonzero(np.all((pv[t0] >= org) & (pvmax[t0] <= top), axis=1))[0]
        tc = tets[cand]
        ok = np.ones(len(tc), bool)
        for q in range(1, 4):
            ok &= np.all((pv[tc[:, q]] >= org) & (pvmax[tc[:, q]] <= top), axis=1)
        got = np.unique(np.sort(keys[tc[ok]], axis=1), axis=0)
        assert len(got) == int(ok.sum())                                              # no duplicate tetrahedra
        assert np.array_equal(got, want)                                              # topology: bit-exact
        pos = np.searchsorted(keys, gkeys)
        assert np.array_equal(keys[pos], gkeys)
        assert np.array_equal(out["lowmin"][pos], r["lowmin"])
        np.testing.assert_allclose(out["verts"][pos], r["pos"] + org.astype(np.float32), rtol=1e-4, atol=1e-4)
        checked += len(want)
    assert checked > 50000


def test_c2_field16384_subblocks_equal_oracle(engine):
    """BASELINE configs[1]: synthetic.field2d(16384), 16 levels by the Linear2DContour rule, exactly as tools/bench_paths.py."""
    import torch
    from contourist_b200 import synthetic
    n = 16384
    f = synthetic.field2d(n)
    mn, mx = float(f.min()), float(f.max())
    levels = [(mx - mn) / 17 * i for i in range(1, 17)]
    c = engine.mt2d_run(f.data_ptr(), levels, shape=(n, n), dtype=np.float32)
    out = engine.mt2d_fetch()
    assert len(out["keys"]) == c.n_segments > 1e6
    keys = out["keys"]                                                               # [S, 2]

    def decode(k):
        lin = (k >> np.uint64(3)).astype(np.int64)
        d = ((k >> np.uint64(1)) & np.uint64(3)).astype(np.int64)
        p = np.stack([lin // n, lin % n], axis=-1)
        return p, p + np.stack([(d >> 1) & 1, d & 1], axis=-1)
    p0, p1 = decode(keys)                                                            # [S, 2, 2]
    lo = np.minimum(p0, p1).min(axis=1)
    hi = np.maximum(p0, p1).max(axis=1)
    rng = np.random.default_rng(2)
    B = 384
    checked = 0
    for org in _blocks(rng, (n, n), (B, B), NBLOCKS):
        org = np.array(org)
        sub = f[org[0]:org[0] + B, org[1]:org[1] + B].cpu().numpy()
        inside = np.all(lo >= org, axis=1) & np.all(hi <= org + B - 1, axis=1)
        for li, z in enumerate(levels):
            r = mt2d.extract_level(sub, z, np.float32)
            sk = r["seg_keys"]                                                       # [s, 2] local keys, rows sorted
            lin = (sk >> np.uint64(3)).astype(np.int64)
            gi, gj = lin // B + org[0], lin % B + org[1]
            gk = ((gi * n + gj).astype(np.uint64) << np.uint64(3)) | (sk & np.uint64(7))
            want = np.unique(np.sort(gk, axis=1), axis=0)
            sel = inside & (out["level"] == li)
            got = np.unique(np.sort(keys[sel], axis=1), axis=0)
            assert int(sel.sum()) == len(got)                                        # no duplicate segments
            assert np.array_equal(got, want)
            # positions of the block's keys
            k_local, first = np.unique(keys[sel].reshape(-1), return_index=True)
            lk = r["keys"]
            llin = (lk >> np.uint64(3)).astype(np.int64)
            gkk = (((llin // B + org[0]) * n + (llin % B + org[1])).astype(np.uint64) << np.uint64(3)) | (lk & np.uint64(7))
            o2 = np.argsort(gkk)
            assert np.array_equal(gkk[o2], k_local)
            np.testing.assert_allclose(out["pos"][sel].reshape(-1, 2)[first], (r["pos"] + org.astype(np.float32))[o2],
                                       rtol=1e-4, atol=1e-4)
            checked += len(want)
    assert checked > 10000


def test_large_single_run_equals_its_slabs(engine):
    """272 x 1024 x 1024 samples in ONE call (8704 tiles of 1024 words: past the size up to which the offset kernel sums
    the tile aggregates itself, so the one-block tile scan runs) against the same planes cut into four z-slabs (each
    below that size): identical keys, positions and triangles."""
    import torch
    from contourist_b200 import engine as E
    from contourist_b200 import sharding, synthetic
    n, n0 = 1024, 272
    f = synthetic.turbulence(n, 0, n0, n_total=n)
    torch.cuda.synchronize()
    flags = E.WANT_KEYS
    c = engine.mt3d_run(f.data_ptr(), 0.0, flags=flags, shape=(n0, n, n), dtype=np.float32)
    ref = engine.mt3d_fetch()
    assert c.n_tris > 1000000
    keys, verts, tris, voff = [], [], [], 0
    plane = n * n * 4
    for (a, b) in sharding.slab_bounds(n0, 4):
        lo, hi, kw = sharding.slab_with_halo(a, b, n0)
        cs = engine.mt3d_run(f.data_ptr() + lo * plane, 0.0, flags=flags, shape=(hi - lo, n, n), dtype=np.float32, **kw)
        o = engine.mt3d_fetch()
        keys.append(o["keys"]); verts.append(o["verts"]); tris.append(o["tris"].astype(np.int64) + voff)
        voff += cs.n_verts
    assert voff == c.n_verts
    assert np.array_equal(np.concatenate(keys), ref["keys"])
    assert np.array_equal(np.concatenate(verts), ref["verts"])
    assert np.array_equal(np.concatenate(tris), ref["tris"].astype(np.int64))
