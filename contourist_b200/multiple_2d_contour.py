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

    def get_contours_dictionary(self):
        grid = self.grid
        levels = sorted(set(float(v) for v in self.values))
        eng = E.default_engine()
        eng.mt2d_run(grid.samples(0), levels, flags=E.GEOM_F64)
        seg = self.segments = eng.mt2d_fetch()
        order = np.argsort(seg["level"], kind="stable")
        bounds = np.searchsorted(seg["level"][order], np.arange(len(levels) + 1))
        out = {}
        for li, value in enumerate(levels):
            sel = order[bounds[li]:bounds[li + 1]]
            grid_contours = triangulated.chain_segments(seg["keys"][sel], seg["pos"][sel])
            out[value] = [(closed, [grid.from_grid_coordinates(p) for p in pts]) for closed, pts in grid_contours]
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
