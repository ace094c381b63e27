"""Multi-level 2D contours -- drop-in for contourist/multiple_2d_contour.py.

Multiple2DContourGrid / Multiple2DContour / Percentile2DContour / Linear2DContour with
get_contours_dictionary() -> {value: [(closed, [points...]), ...]} (multiple_2d_contour.py:17-30).
All levels are extracted by ONE pass of the GPU over the field (the reference runs one Python search per level).
"""
import numpy as np

from . import engine as E
from . import field2d
from . import triangulated


class Multiple2DContourGrid(object):

    def __init__(self, function_grid, values, segment_endpoints=()):
        self.grid = function_grid
        self.values = list(sorted(values))
        self.segment_end_points = segment_endpoints
        self.value_to_endpoints = None
        self.value_to_contour_sequences = None
        self.segments = None

    keep_segments = False             # True: fetch the segment soup (self.segments) and chain it on the host
    MAX_LEVELS_PER_PASS = 64          # ctr_mt2d_params.nlevels: one GPU pass classifies against up to 64 levels

    def get_contours_dictionary(self):
        """{value: [(closed, points[k, 2]), ...]} (multiple_2d_contour.py:17-30).  One GPU pass per 64 distinct levels;
        the engine applies grid_field.from_grid_coordinates (x * delta + mins) itself, so the points come back in world
        coordinates as arrays (iterate them like the reference's lists of points)."""
        grid = self.grid
        levels = sorted(set(float(v) for v in self.values))
        eng = E.default_engine()
        field = grid.samples(0)
        out = {}
        self.segments = []
        for c0 in range(0, len(levels), self.MAX_LEVELS_PER_PASS):
            chunk = levels[c0:c0 + self.MAX_LEVELS_PER_PASS]
            eng.mt2d_run(field, chunk, origin=tuple(grid.mins), delta=tuple(grid.delta), flags=E.GEOM_F64)
            polys = None if self.keep_segments else eng.mt2d_polylines()       # chained on the device
            if polys is not None:
                for li, value in enumerate(chunk):
                    out[value] = polys[li]
                continue
            # some key lies on more than two segments (a sample exactly on a level): the fixed-order walk decides
            seg = eng.mt2d_fetch()
            self.segments.append(seg)
            order = np.argsort(seg["level"], kind="stable")
            bounds = np.searchsorted(seg["level"][order], np.arange(len(chunk) + 1))
            for li, value in enumerate(chunk):
                sel = order[bounds[li]:bounds[li + 1]]
                out[value] = triangulated.chain_segments(seg["keys"][sel], seg["pos"][sel])
        self.segments = self.segments[0] if len(self.segments) == 1 else (self.segments or None)
        self.value_to_contour_sequences = {v: out[float(v)] for v in self.values}
        return self.value_to_contour_sequences


class Multiple2DContour(Multiple2DContourGrid):

    def __init__(self, xmin, ymin, xmax, ymax, dx, dy, function, values, segment_endpoints=()):
        function_grid = field2d.Function2DGrid(xmin, ymin, xmax, ymax, dx, dy, function)
        Multiple2DContourGrid.__init__(self, function_grid, values, segment_endpoints)


class Percentile2DContour(Multiple2DContourGrid):

    def __init__(self, xmin, ymin, xmax, ymax, dx, dy, function, breakpoints=10, segment_endpoints=()):
        function_grid = field2d.Function2DGrid(xmin, ymin, xmax, ymax, dx, dy, function)
        self.function_grid = function_grid
        values = self.values = self.get_values(breakpoints)
        Multiple2DContourGrid.__init__(self, function_grid, values, segment_endpoints)

    def get_values(self, breakpoints):
        "Sorted samples at stride n/breakpoints (multiple_2d_contour.py:91-98)."
        samples = np.sort(self.function_grid.samples(0).reshape(-1))
        (nsamples,) = samples.shape
        skip = int(nsamples / breakpoints)
        return [samples[index] for index in range(skip, nsamples, skip)]


class Linear2DContour(Percentile2DContour):

    def get_values(self, breakpoints):
        "(max - min)/breakpoints * i for i in 1..breakpoints-1 -- not shifted by the minimum (multiple_2d_contour.py:100-108)."
        eng = E.default_engine()
        c = eng.mt2d_run(self.function_grid.samples(0), [0.0], flags=E.WANT_MINMAX | E.NO_GEOMETRY)
        offset = (c.fmax - c.fmin) * (1.0 / breakpoints)
        return [offset * i for i in range(1, breakpoints)]
