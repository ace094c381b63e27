"""find_contour_crossing_grid_segments on the engine (grid_field.py:64-84)."""
import numpy as np

from . import engine as E


def crossing_segments(grid, value, skip=1):
    d = grid.dimension
    if skip != 1:
        raise NotImplementedError("skip != 1: the GPU engine always scans every grid vertex (the reference's skip only "
                                  "thins the seed search, grid_field.py:64,71)")
    if d == 3:
        arr = grid.samples(1)
        eng = E.default_engine()
        c = eng.mt3d_run(arr, value, flags=E.WANT_KEYS | E.WANT_MINMAX | E.GEOM_F64)
        out = eng.mt3d_fetch(verts=False, tris=False)
        keys = out["keys"]
        n = arr.shape
        lin = (keys >> np.uint64(3)).astype(np.int64)
        dd = (keys & np.uint64(7)).astype(np.int64)
        p = np.stack([lin // (n[1] * n[2]), (lin // n[2]) % n[1], lin % n[2]], axis=1)
        q = p + np.stack([(dd >> 2) & 1, (dd >> 1) & 1, dd & 1], axis=1)
    elif d == 4:
        arr = grid.samples(1)
        eng = E.default_engine()
        c = eng.mp4d_run(arr, value, flags=E.WANT_KEYS | E.WANT_MINMAX | E.GEOM_F64)
        out = eng.mp4d_fetch(verts=False, tets=False, morph=False)
        keys = out["keys"]
        n = arr.shape
        lin = (keys >> np.uint64(4)).astype(np.int64)
        dd = (keys & np.uint64(15)).astype(np.int64)
        p = np.stack([lin // (n[1] * n[2] * n[3]), (lin // (n[2] * n[3])) % n[1], (lin // n[3]) % n[2], lin % n[3]], axis=1)
        q = p + np.stack([(dd >> 3) & 1, (dd >> 2) & 1, (dd >> 1) & 1, dd & 1], axis=1)
    else:
        raise NotImplementedError("find_contour_crossing_grid_segments: 3D and 4D grids (2D grids are searched by "
                                  "Multiple2DContour / Grid2DContour on the GPU)")
    # the reference visits v0 in range(0, N)^d only, and needs (f0-v)*(f1-v) < 0 strictly
    inr = np.all(p < (np.array(arr.shape) - 1), axis=1)
    f0 = arr[tuple(p.T)].astype(float) - value
    f1 = arr[tuple(q.T)].astype(float) - value
    strict = (f0 * f1) < 0
    sel = inr & strict
    assert int(sel.sum()) == int(c.n_crossings), "engine crossing count disagrees with its own edge keys"
    return (c.fmax, c.fmin, SegmentList(p[sel], q[sel]))


class SegmentList(object):
    """Sequence of (v0, v1) integer grid-vertex pairs backed by two arrays (the reference builds a Python list)."""

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
