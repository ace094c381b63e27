"""3D isosurfaces by marching tetrahedra -- drop-in for contourist/tetrahedral.py on the CUDA engine.

Same public names and signatures as the reference (tetrahedral.py:50-107,514-621): TriangulatedIsosurfaces,
Delta3DContour, Grid3DContour / GridContour3d with search_for_endpoints(), get_points_and_triangles(),
extract_surface_geometry().  Differences, all documented in DESIGN.md section 2:
  * `function` may be a numpy array of samples (N+1 per axis for N = grid_dimensions) besides a callable;
  * extraction is a FULL SCAN on the GPU: segment_endpoints only matter as "non-empty or not" (the reference
    at HEAD cannot take 3D seeds through this facade either, triangulated.py:96);
  * linear_interpolate=False (the callable evaluated off-grid) is not available: NotImplementedError;
  * points / triangles come back as numpy arrays (iterate / index them like the reference's lists);
  * the reference's serial mesh post-processing (quantize, tiny, clean, global orientation) is replaced by the
    engine's deterministic indexed mesh, wound towards the high side; `reference_orientation=True` adds the
    reference's per-component outward flip.
"""
import numpy as np

from . import engine as E
from . import grid_field
from . import surface_geometry

A = (0, 0, 0)
B = (0, 0, 1)
C = (0, 1, 0)
D = (0, 1, 1)
E_ = (1, 0, 0)
F = (1, 0, 1)
G = (1, 1, 0)
H = (1, 1, 1)
CUBE = np.array([A, B, C, D, E_, F, G, H], dtype=int)
TETRAHEDRA = np.array([[A, H, B, D], [A, H, D, C], [A, H, C, G], [A, H, G, E_], [A, H, E_, F], [A, H, F, B]], dtype=int)


class GridContour3d(object):
    """Grid-coordinate engine front end (reference: GridContour + GridContour3d, tetrahedral.py:109-621)."""

    flatten = False
    smooth = None
    minimum_ratio = 0.05
    minimum_extent = None
    geometry_dtype = np.float64
    reference_orientation = False
    want_normals = False

    def __init__(self, corner, function, value, segment_endpoints, linear_interpolate=True, callback=None,
                 origin=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0)):
        self.corner = np.array(corner, dtype=int)
        (self.dimension,) = self.corner.shape
        self.sanity_check()
        if not linear_interpolate:
            raise NotImplementedError("linear_interpolate=False evaluates the callable off-grid "
                                      "(tetrahedral.py:488-505); the array engine cannot")
        self.linear_interpolate = linear_interpolate
        self.end_points = segment_endpoints
        self.f = function
        self.value = value
        self.callback = callback
        self.origin = origin
        self.delta = delta
        self.normals = None
        self.counts = None

    def sanity_check(self):
        assert self.dimension == 3

    def _field(self):
        if isinstance(self.f, np.ndarray):
            want = tuple(int(c) + 1 for c in self.corner)
            if tuple(self.f.shape) != want:
                raise ValueError("sample array has shape %r, corner %r needs %r" % (self.f.shape, tuple(self.corner), want))
            return self.f
        n = [int(c) + 1 for c in self.corner]
        g = grid_field.FunctionGrid([0, 0, 0], [c for c in self.corner], [1, 1, 1], self.f)
        assert tuple(g.grid_dimensions) == tuple(n), (tuple(g.grid_dimensions), n)
        return g.samples(0)

    def get_points_and_triangles(self, clean=True):
        if self.flatten or self.smooth:
            raise NotImplementedError("flatten / smooth (LP-based decimation, tetrahedral.py:217-351) are out of scope")
        eng = E.default_engine()
        flags = (E.GEOM_F64 if np.dtype(self.geometry_dtype) == np.float64 else 0) | (E.WANT_NORMALS if self.want_normals else 0)
        self.counts = eng.mt3d_run(self._field(), self.value, origin=self.origin, delta=self.delta, flags=flags)
        if self.reference_orientation:
            # surface_geometry.py:52-140 on the device mesh: one keep / reverse decision per edge-connected component
            self.components, self.flipped = eng.mt3d_orient_reference()
        out = eng.mt3d_fetch()
        self.normals = out["normals"]
        points, triangles = out["verts"], out["tris"]
        if self.callback:
            self.callback(self)
        return (points, triangles)

    def extract_surface_geometry(self, clean=True):
        points, triangles = self.get_points_and_triangles(clean)
        geometry = surface_geometry.SurfaceGeometry(points, triangles)
        if clean:
            geometry.clean_triangles()
        geometry.orient_triangles()
        return geometry


def Grid3DContour(horizontal_n, vertical_m, forward_l, function, value, segment_endpoints,
                  linear_interpolate=True, callback=None):
    return GridContour3d((horizontal_n, vertical_m, forward_l), function, value, segment_endpoints, linear_interpolate,
                         callback)


class Delta3DContour(object):
    """World-coordinate driver (reference: triangulated.ContourGrid + tetrahedral.Delta3DContour)."""

    linear_interpolate = True
    flatten = False
    minimum_ratio = None
    minimum_extent = None
    smooth = None

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
        maker = GridContour3d(tuple(int(n) for n in grid.grid_dimensions), grid.samples(1), self.value, grid_endpoints,
                              linear_interpolate=self.linear_interpolate, origin=tuple(grid.mins), delta=tuple(grid.delta))
        maker.flatten = self.flatten
        maker.smooth = self.smooth
        return maker

    def search_for_endpoints(self, skip=1):
        (maxf, minf, grid_endpoints) = self.grid.find_contour_crossing_grid_segments(self.value, skip)
        self.grid_endpoints = grid_endpoints
        self.contour_maker = self.get_contour_maker(grid_endpoints)

    def get_points_and_triangles(self):
        # the engine applies grid_field.from_grid_coordinates (x*delta + mins) on the device
        return self.contour_maker.get_points_and_triangles()


class TriangulatedIsosurfaces(Delta3DContour):

    def __init__(self, mins, maxes, delta, function, value, segment_endpoints,
                 linear_interpolate=True, flatten=False, minimum_ratio=None, minimum_extent=None, smooth=None):
        self.flatten = flatten
        self.smooth = smooth
        if minimum_ratio is not None:
            self.minimum_ratio = minimum_ratio
        if minimum_extent is not None:
            self.minimum_extent = minimum_extent
        grid = grid_field.FunctionGrid(mins, maxes, delta, function)
        Delta3DContour.__init__(self, grid, value, segment_endpoints, linear_interpolate=linear_interpolate)
