"""Grid context for a scalar field over n dimensions -- drop-in for contourist/grid_field.py.

Same class, constructor and method names as the reference (grid_field.py:8-118).  Extensions:
  * `function` may be a numpy array of samples (index [i, j, ...] = grid point (i, j, ...)) instead of a callable;
  * `samples(extra)` materialises the callable once into the dense array the CUDA engine consumes, trying a
    vectorised call first and falling back to the reference's per-point Python loop (grid_field.py:34-43);
  * `find_contour_crossing_grid_segments` runs on the GPU (the reference scans every vertex in Python).
"""
import numpy as np


class FunctionGrid(object):

    def __init__(self, mins, maxes, delta, function, materialize=False, cache=False):
        self.mins = np.array(mins, dtype=float)
        shape = self.mins.shape
        self.maxes = np.zeros(shape, dtype=float)
        self.maxes[:] = maxes
        self.delta = np.zeros(shape, dtype=float)
        self.delta[:] = delta
        (self.dimension,) = shape
        self.f = function
        self.cached = cache
        self.materialize = materialize
        self.materialized_array = None
        self.cache = {}
        self.grid_dimensions = self.to_grid_vertex(self.maxes) + 1          # grid_field.py:26-27 (truncation)
        assert np.all(self.grid_dimensions >= 2), "grid must have dimensions greater than 2"
        self._samples = {}
        if isinstance(function, np.ndarray):
            assert function.ndim == self.dimension, "sample array rank must equal the grid dimension"
        if materialize:
            assert not cache, "do not cache and materialize at the same time."
            self.materialize_array()

    # ---- coordinates (grid_field.py:45-50,86-93)
    def to_grid_coordinates(self, xypoint):
        return (xypoint - self.mins) / self.delta

    def to_grid_vertex(self, xypoint):
        return np.array(self.to_grid_coordinates(xypoint), dtype=int)

    def from_grid_coordinates(self, xygrid):
        return (np.array(xygrid, dtype=float) * self.delta) + self.mins

    def on_grid(self, grid_vertex):
        return np.all(grid_vertex >= 0) and np.all(grid_vertex <= self.grid_dimensions)

    def surrounding_vertices(self, xypoint, skip=1, grid_vertex=False):
        vertex0 = xypoint if grid_vertex else self.to_grid_vertex(xypoint)
        for index in range(2 ** self.dimension):
            offset = np.array([((index >> s) & 1) * skip for s in range(self.dimension)], dtype=int)
            yield vertex0 + offset

    # ---- sampling
    def grid_function(self, *xy_grid):
        """f at grid coordinates (grid_field.py:95-118): materialised array, then cache, then the callable."""
        xy_grid = tuple(xy_grid)
        all_ints = all(isinstance(x, (int, np.integer)) for x in xy_grid)
        if isinstance(self.f, np.ndarray):
            idx = tuple(min(max(int(x), 0), n - 1) for x, n in zip(xy_grid, self.f.shape))
            return self.f[idx]
        m = self.materialized_array
        if m is not None and all_ints:
            try:
                return m[tuple(int(x) for x in xy_grid)]
            except IndexError:
                pass
        if self.cached and all_ints and xy_grid in self.cache:
            return self.cache[xy_grid]
        result = self.f(*self.from_grid_coordinates(xy_grid))
        if self.cached and all_ints:
            self.cache[xy_grid] = result
        return result

    def samples(self, extra=0, dtype=None):
        """Dense sample array of shape grid_dimensions + extra (3D/4D voxel engines need extra=1: voxel N-1 reads
        sample N, tetrahedral.py:465-469).  Vectorised evaluation when the callable broadcasts, else the
        reference's per-point loop."""
        key = (int(extra), None if dtype is None else np.dtype(dtype).str)
        if key in self._samples:
            return self._samples[key]
        shape = tuple(int(n) + int(extra) for n in self.grid_dimensions)
        if isinstance(self.f, np.ndarray):
            if tuple(self.f.shape) != shape:
                raise ValueError("sample array has shape %r, the grid needs %r (grid_dimensions + %d)"
                                 % (tuple(self.f.shape), shape, extra))
            arr = self.f
        else:
            arr = None
            axes = [np.arange(n, dtype=float) * d + m for n, d, m in zip(shape, self.delta, self.mins)]
            try:
                mesh = np.meshgrid(*axes, indexing="ij", sparse=True)
                cand = np.asarray(self.f(*mesh), dtype=float)
                cand = np.broadcast_to(cand, shape) if cand.shape != shape else cand
                # a callable that silently mis-broadcasts is caught by probing a few points
                rng = np.random.default_rng(0)
                ok = True
                for _ in range(4):
                    idx = tuple(int(rng.integers(0, n)) for n in shape)
                    ref = float(self.f(*[a[i] for a, i in zip(axes, idx)]))
                    if not (cand[idx] == ref or (np.isnan(cand[idx]) and np.isnan(ref))):
                        ok = False
                if ok:
                    arr = np.ascontiguousarray(cand)
            except Exception:
                arr = None
            if arr is None:
                arr = np.zeros(shape, dtype=float)
                for index in iter_indices(shape):
                    arr[index] = self.f(*[a[i] for a, i in zip(axes, index)])
        if dtype is not None:
            arr = np.ascontiguousarray(arr, dtype=dtype)
        self._samples[key] = arr
        return arr

    def materialize_array(self):
        self.materialized_array = np.array(self.samples(0), dtype=float)
        return self.materialized_array

    def find_contour_crossing_grid_segments(self, value, skip=1):
        """(maxf, minf, [(v0, v1), ...]) for every strictly crossing grid segment v0 -> v0 + {0,1}^d
        (grid_field.py:64-84).  Runs on the GPU for 3D and 4D grids with skip == 1; the segment list is
        derived from the engine's edge keys."""
        from . import _segments
        return _segments.crossing_segments(self, value, skip)


def iter_indices(shape, skip=1):
    "All index tuples of an array of the given shape (grid_field.py:120-137)."
    if len(shape) == 0:
        yield ()
        return
    if len(shape) == 1:
        for i in range(0, shape[0], skip):
            yield (i,)
        return
    tails = list(iter_indices(shape[1:], skip))
    for i in range(0, shape[0], skip):
        for tail in tails:
            yield (i,) + tail
