"""find_contour_crossing_grid_segments on the engine (grid_field.py:64-84)."""
import numpy as np

from . import engine as E


def _scan_samples(arr, value, skip):
    """grid_field.py:64-84 on a sample array, vectorised: vertex0 over range(0, N, skip) per axis (N = samples - 1),
    vertex1 = vertex0 + skip * offset for every non-zero offset in {0,1}^d, strict crossing (f0-v)*(f1-v) < 0.
    Segments whose far end lies outside the array (the reference would call the function off-grid there) are left
    out.  Host code: it thins the seed search, reads (N/skip)^d samples, and is exact for any field."""
    d = arr.ndim
    n = [s - 1 for s in arr.shape]
    starts = [np.arange(0, k, skip) for k in n]
    ps, qs = [], []
    fmax, fmin = -np.inf, np.inf
    mesh = np.meshgrid(*starts, indexing="ij") if all(len(s) for s in starts) else None
    if mesh is None:
        return np.nan, np.nan, np.zeros((0, d), np.int64), np.zeros((0, d), np.int64)
    p0 = np.stack([m.reshape(-1) for m in mesh], axis=1)
    f0 = arr[tuple(p0.T)].astype(np.float64)
    fmax, fmin = max(fmax, f0.max()), min(fmin, f0.min())
    for index in range(1, 2 ** d):
        off = np.array([((index >> s) & 1) * skip for s in range(d)], dtype=np.int64)      # grid_field.py:59-61
        p1 = p0 + off
        ok = np.all(p1 < np.array(arr.shape), axis=1)
        a, b = p0[ok], p1[ok]
        fa, fb = f0[ok], arr[tuple(b.T)].astype(np.float64)
        if len(fb):
            fmax, fmin = max(fmax, fb.max()), min(fmin, fb.min())
        sel = (fa - value) * (fb - value) < 0
        ps.append(a[sel])
        qs.append(b[sel])
    return fmax, fmin, np.concatenate(ps), np.concatenate(qs)


def crossing_segments(grid, value, skip=1):
    d = grid.dimension
    if d not in (3, 4):
        raise NotImplementedError("find_contour_crossing_grid_segments: 3D and 4D grids (2D grids are searched by "
                                  "Multiple2DContour / Grid2DContour on the GPU)")
    arr = grid.samples(1)
    if skip != 1:
        fmax, fmin, p, q = _scan_samples(arr, value, int(skip))
        return (fmax, fmin, SegmentList(p, q))
    eng = E.default_engine()
    if d == 3:
        # the scan is stages 1-2 of the extraction itself: run it in the form get_points_and_triangles() wants (grid
        # coordinates, fp64) and leave the mesh on the device, so that the extraction that follows only fetches
        flags = E.WANT_KEYS | E.WANT_MINMAX | E.GEOM_F64
        c = eng.mt3d_run(arr, value, flags=flags)
        run = (eng, eng.run_serial, flags, c)
        out = eng.mt3d_fetch(verts=False, tris=False)
        keys = out["keys"]
        n = arr.shape
        lin = (keys >> np.uint64(3)).astype(np.int64)
        dd = (keys & np.uint64(7)).astype(np.int64)
        p = np.stack([lin // (n[1] * n[2]), (lin // n[2]) % n[1], lin % n[2]], axis=1)
        q = p + np.stack([(dd >> 2) & 1, (dd >> 1) & 1, dd & 1], axis=1)
    else:
        c = eng.mp4d_run(arr, value, flags=E.WANT_KEYS | E.WANT_MINMAX | E.GEOM_F64)
        run = None
        out = eng.mp4d_fetch(verts=False, tets=False, morph=False)
        keys = out["keys"]
        n = arr.shape
        lin = (keys >> np.uint64(4)).astype(np.int64)
        dd = (keys & np.uint64(15)).astype(np.int64)
        p = np.stack([lin // (n[1] * n[2] * n[3]), (lin // (n[2] * n[3])) % n[1], (lin // n[3]) % n[2], lin % n[3]], axis=1)
        q = p + np.stack([(dd >> 3) & 1, (dd >> 2) & 1, (dd >> 1) & 1, dd & 1], axis=1)
    # the reference visits v0 in range(0, N)^d only, and needs (f0-v)*(f1-v) < 0 strictly
    inr = np.all(p < (np.array(arr.shape) - 1), axis=1)
    f0 = arr[tuple(p.T)].astype(float) - value
    f1 = arr[tuple(q.T)].astype(float) - value
    sel = inr & ((f0 * f1) < 0)
    p, q = p[sel], q[sel]
    if len(p) != int(c.n_crossings):
        # Edge keys exist for the edges of emitted triangles only.  A strictly crossing segment all of whose tetrahedra
        # are skipped by the np.allclose rule (tetrahedral.py:576; samples within 1e-5 relative of the isovalue on both
        # sides) has no key, but the engine's exact count (k_count_b / count_word_exact) includes it: list those from
        # the samples.
        _, _, p, q = _scan_samples(arr, value, 1)
        assert len(p) == int(c.n_crossings), "engine crossing count disagrees with the samples"
    seg = SegmentList(p, q)
    seg.engine_run = run
    return (c.fmax, c.fmin, seg)


class SegmentList(object):
    """Sequence of (v0, v1) integer grid-vertex pairs backed by two arrays (the reference builds a Python list)."""

    engine_run = None          # (engine, run_serial, flags, counts) of the extraction that produced the list, if any

    def __init__(self, p, q):
        self.p = p
        self.q = q

    def __len__(self):
        return len(self.p)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SegmentList(self.p[i], self.q[i])
        return (self.p[i], self.q[i])

    def __iter__(self):
        for a, b in zip(self.p, self.q):
            yield (a, b)
