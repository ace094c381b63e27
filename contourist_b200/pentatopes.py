"""4D isosurfaces by marching pentatopes / morphing 3D surfaces -- drop-in for contourist/pentatopes.py.

MorphingIsoSurfaces / Delta4DContour / GridContour4D with search_for_endpoints(), find_tetrahedra(),
collect_morph_triangles(), to_json() (pentatopes.py:42-125,314-368).  The heavy work (pentatope classification,
tetrahedra, time binning, instant/tiny filtering, slicing every tetrahedron into morph triangles) runs on the GPU;
this module only numbers the unique segments / triangles the way the reference's dicts do and packs the result.
"""
import itertools

import numpy as np

from . import engine as E
from . import grid_field
from . import morph_geometry


def _pentatope_tiles():
    out = []
    for permutation in itertools.permutations(range(4)):
        vertex = [0, 0, 0, 0]
        tile = [vertex[:]]
        for index in permutation:
            vertex[index] = 1
            tile.append(vertex[:])
        out.append(tile)
    return out


PENTATOPES = np.array(_pentatope_tiles(), dtype=int)
HYPERCUBE = np.array([(i, j, k, l) for i in (0, 1) for j in (0, 1) for k in (0, 1) for l in (0, 1)], dtype=int)


class GridContour4D(object):

    minimum_ratio = 0.05
    flatten = False
    smooth = None

    def __init__(self, corner, function, value, segment_endpoints, linear_interpolate=True, callback=None):
        self.corner = np.array(corner, dtype=int)
        (self.dimension,) = self.corner.shape
        self.sanity_check()
        if not linear_interpolate:
            raise NotImplementedError("linear_interpolate=False is not available on the array engine")
        self.f = function
        self.value = value
        self.end_points = segment_endpoints
        self.callback = callback
        self.counts = None
        self._out = None

    def sanity_check(self):
        assert self.dimension == 4, "dimension should be 4 " + repr(self.dimension)

    def _field(self):
        if isinstance(self.f, np.ndarray):
            want = tuple(int(c) + 1 for c in self.corner)
            if tuple(self.f.shape) != want:
                raise ValueError("sample array has shape %r, corner %r needs %r" % (self.f.shape, tuple(self.corner), want))
            return self.f
        g = grid_field.FunctionGrid([0] * 4, list(self.corner), [1] * 4, self.f)
        return g.samples(0)

    def find_tetrahedra(self):
        "pentatopes.py:101-125 (not flatten, not smooth): raw tetrahedra, bin_times, drop_instant, remove_tiny."
        if self.flatten or self.smooth:
            raise NotImplementedError("flatten / smooth are out of scope (tetrahedral.py:217-351)")
        eng = E.default_engine()
        self.counts = eng.mp4d_run(self._field(), self.value, flags=E.GEOM_F64 | E.MORPH)
        self._out = eng.mp4d_fetch()
        self.dropped_simplices = int((self._out["keep"] == 0).sum())
        return self._out

    def collect_morph_triangles(self, epsilon=1e-7):
        "pentatopes.py:314-368: unique segments, unique triangles over them, time-aware orientation."
        if self._out is None:
            self.find_tetrahedra()
        v4 = self._out["morph_verts"]
        mt = self._out["morph_tris"].astype(np.int64)                  # [K, 3, 2] (low-t id, high-t id), with duplicates
        nv = len(v4) + 1
        skey = np.minimum(mt[:, :, 0], mt[:, :, 1]) * nv + np.maximum(mt[:, :, 0], mt[:, :, 1])
        useg, sid = np.unique(skey.reshape(-1), return_inverse=True)
        tri = np.sort(sid.reshape(-1, 3), axis=1)
        tri = np.unique(tri, axis=0) if len(tri) else tri.reshape(0, 3)
        segments = np.stack([useg // nv, useg % nv], axis=1)
        result = morph_geometry.MorphTriangles(v4, segments, tri)
        result.orient_triangles()
        return result


class Delta4DContour(object):

    flatten = False
    minimum_ratio = None
    minimum_extent = None
    smooth = None
    linear_interpolate = True

    def __init__(self, function_grid, value, segment_endpoints=None, linear_interpolate=True):
        self.linear_interpolate = linear_interpolate
        self.grid = function_grid
        self.value = value
        self.segment_endpoints = segment_endpoints
        self.grid_endpoints = None
        self.contour_maker = self.get_contour_maker(None)

    def get_contour_maker(self, grid_endpoints):
        grid = self.grid
        self.grid_endpoints = grid_endpoints
        maker = GridContour4D(tuple(int(n) for n in grid.grid_dimensions), grid.samples(1), self.value, grid_endpoints,
                              linear_interpolate=self.linear_interpolate)
        maker.flatten = self.flatten
        maker.smooth = self.smooth
        return maker

    def search_for_endpoints(self, skip=1):
        (maxf, minf, grid_endpoints) = self.grid.find_contour_crossing_grid_segments(self.value, skip)
        self.grid_endpoints = grid_endpoints
        self.contour_maker = self.get_contour_maker(grid_endpoints)

    def collect_morph_triangles(self):
        self.contour_maker.find_tetrahedra()
        return self.contour_maker.collect_morph_triangles().from_grid_coordinates(self.grid)


class MorphingIsoSurfaces(Delta4DContour):

    def __init__(self, mins, maxes, delta, function, value, segment_endpoints,
                 linear_interpolate=True, flatten=False, minimum_ratio=None, minimum_extent=None, smooth=None):
        self.flatten = flatten
        self.smooth = smooth
        if minimum_ratio is not None:
            self.minimum_ratio = minimum_ratio
        if minimum_extent is not None:
            self.minimum_extent = minimum_extent
        grid = grid_field.FunctionGrid(mins, maxes, delta, function)
        Delta4DContour.__init__(self, grid, value, segment_endpoints, linear_interpolate=linear_interpolate)

    def to_json(self):
        morph_triangles = self.collect_morph_triangles()
        return morph_triangles.to_json(min_value=self.grid.mins[-1], max_value=self.grid.maxes[-1])
