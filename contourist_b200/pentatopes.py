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
OFFSETS4D = np.array([(i, j, k, l) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) for l in (-1, 0, 1)
                      if i != 0 or j != 0 or k != 0 or l != 0], dtype=int)


def border_hypervoxels(samples, value):
    "tetrahedral.py:383-394 border_voxel for every hypervoxel of a 4D sample array: not np.allclose(value, f), min <= value <= max"
    s = np.asarray(samples, dtype=np.float64)
    n = [k - 1 for k in s.shape]
    mn = mx = None
    allnear = None
    for c in HYPERCUBE:
        v = s[tuple(slice(int(o), int(o) + k) for o, k in zip(c, n))]
        mn = v.copy() if mn is None else np.minimum(mn, v)
        mx = v.copy() if mx is None else np.maximum(mx, v)
        near = np.abs(value - v) <= 1e-8 + 1e-5 * np.abs(v)
        allnear = near if allnear is None else (allnear & near)
    return (~allnear) & (mn <= value) & (mx >= value)


def initial_hypervoxels(samples, value, end_points, border):
    """Start hypervoxels of the reference's tracker for seed segments (tetrahedral.py:396-441 with the 80 offsets of
    pentatopes.py:32-39): bisection of every seed down to adjacent points, then each of the two points -- or the first
    of its 80 neighbours in OFFSETS4D order -- that is a border hypervoxel, with the reference's `visited` set."""
    samples = np.asarray(samples)

    def f(p):
        if np.any(np.asarray(p) < 0) or np.any(np.asarray(p) >= np.array(samples.shape)):
            raise ValueError("seed point %r outside the sample array of shape %r" % (tuple(int(x) for x in p), samples.shape))
        return float(samples[tuple(int(x) for x in p)])

    def is_border(p):
        return bool(all(0 <= int(p[a]) < border.shape[a] for a in range(4)) and border[tuple(int(x) for x in p)])

    visited, found = set(), set()
    for low_point, high_point in np.array(end_points, dtype=int).reshape(-1, 2, 4):
        low_value, high_value = f(low_point), f(high_point)
        if low_value > value or high_value < value:
            (low_point, low_value, high_point, high_value) = (high_point, high_value, low_point, low_value)
        assert low_value <= value and high_value >= value, \
            "Bad end points " + repr((tuple(low_point), low_value, tuple(high_point), high_value, value))
        while np.any(np.abs(low_point - high_point) > 1):
            mid_point = (low_point + high_point) // 2
            if f(mid_point) < value:
                low_point = mid_point
            else:
                high_point = mid_point
        for point in (low_point, high_point):
            tpoint = tuple(int(x) for x in point)
            if tpoint in visited:
                continue
            visited.add(tpoint)
            if is_border(point):
                found.add(tpoint)
                continue
            for offset_point in OFFSETS4D + point.reshape(1, 4):
                toffset = tuple(int(x) for x in offset_point)
                if toffset in visited:
                    continue
                visited.add(toffset)
                if is_border(offset_point):
                    found.add(toffset)
                    break
    return sorted(found)


class GridContour4D(object):

    minimum_ratio = 0.05
    flatten = False
    smooth = None
    full_scan = False               # set by search_for_endpoints(): the seeds are every crossing segment

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
        field = self._field()
        seeded = self.end_points is not None and not self.full_scan
        self.counts = eng.mp4d_run(field, self.value, flags=E.GEOM_F64 | E.MORPH | (E.WANT_KEYS if seeded else 0))
        self._out = eng.mp4d_fetch()
        if seeded:
            self._select_seeded(field)
        self.dropped_simplices = int((self._out["keep"] == 0).sum())
        return self._out

    def _select_seeded(self, field):
        """Explicit seed segments (pentatopes.py:92-106 through tetrahedral.py:396-469): keep the tetrahedra of the
        80-connected components of border hypervoxels that contain a start hypervoxel.  The engine has extracted the full
        scan; this restricts its output on the host (a labelling of the border mask), like the 2D seed filter.  A
        tetrahedron's hypervoxel is the component-wise minimum over the edges of its four vertices (they touch all five
        corners of its pentatope, the hypervoxel origin among them); components share no vertex, so a morph triangle
        belongs to the selection iff its first vertex does."""
        from scipy import ndimage
        o = self._out
        border = border_hypervoxels(field, float(self.value))
        self.start_voxels = initial_hypervoxels(field, float(self.value), self.end_points, border) if len(self.end_points) else []
        lab, _ = ndimage.label(border, structure=np.ones((3, 3, 3, 3), dtype=bool))
        want = sorted(set(int(lab[v]) for v in self.start_voxels))
        n = field.shape
        lin = (o["keys"] >> np.uint64(4)).astype(np.int64)
        pmin = np.stack([lin // (n[1] * n[2] * n[3]), (lin // (n[2] * n[3])) % n[1], (lin // n[3]) % n[2], lin % n[3]], axis=1)
        tets = o["tets"].astype(np.int64)
        vox = pmin[tets].min(axis=1) if len(tets) else np.zeros((0, 4), np.int64)
        keep_t = np.isin(lab[tuple(vox.T)], want) if (len(tets) and want) else np.zeros(len(tets), bool)
        used = np.zeros(len(o["keys"]), bool)
        used[tets[keep_t].reshape(-1)] = True
        remap = np.cumsum(used) - 1
        mt = o["morph_tris"].astype(np.int64)
        keep_m = used[mt[:, 0, 0]] if len(mt) else np.zeros(0, bool)
        self._out = dict(o, verts=o["verts"][used], keys=o["keys"][used], lowmin=o["lowmin"][used], morph_verts=o["morph_verts"][used],
                         tets=remap[tets[keep_t]].astype(np.int32), keep=o["keep"][keep_t],
                         morph_tris=remap[mt[keep_m]].astype(np.int32))
        self.selected_voxels = np.isin(lab, want) & border if want else np.zeros_like(border)

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


    # ---- the legacy wire format (pentatopes.py:370-444; misc/morph_sequence.js)
    def iterate_morph_geometry(self):
        "One MorphGeometry per interval between consecutive distinct vertex times (pentatopes.py:370-413)."
        if self._out is None:
            self.find_tetrahedra()
        tets = self._out["tets"][self._out["keep"].astype(bool)]
        return morph_geometry.morph_sequence(self._out["morph_verts"], tets)

    def json_stats(self, morphs=None):
        morphs = list(self.iterate_morph_geometry()) if morphs is None else morphs
        return (morphs, morphs[0].min_value, morphs[-1].max_value)

    def json_data(self, morphs=None):
        (morphs, min_value, max_value) = self.json_stats(morphs)
        return {"min_value": min_value, "max_value": max_value, "morph_descriptions": [m.json_data() for m in morphs]}

    def to_json0(self, morphs=None):
        import json
        return json.dumps(self.json_data(morphs), indent=4, default=lambda x: x.item() if hasattr(x, "item") else list(x))

    def to_json(self, morphs=None):
        '"Sequence of morphing triangularizations." (pentatopes.py:430-444)'
        (morphs, _, _) = self.json_stats(morphs)
        return morph_geometry.morph_sequence_json(morphs)


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
        maker.full_scan = True      # this driver only ever passes the complete seed set (grid search)
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
