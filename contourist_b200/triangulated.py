"""2D contours on the (1,1)-diagonal triangulated grid -- drop-in for contourist/triangulated.py.

Grid2DContour / DxDy2DContourGrid / DxDy2DContour with get_contour_sequences(), adjacent_pairs,
contour_sequences_to_svg (triangulated.py:10-146,148-378).  Keys, positions and the segments that link them come
from the GPU (mt2d.cu); this module chains the segment soup into polylines (triangulated.py:236-293) with a
deterministic start/direction rule instead of the reference's set-iteration order.
"""
import numpy as np

from . import engine as E
from . import field2d

adjacent_offsets = [(0, 1), (1, 1), (1, 0), (0, -1), (-1, -1), (-1, 0)]
adjacency_array = np.array(adjacent_offsets, dtype=int)

SVG_TEMPLATE = """
<svg height="%s" width="%s" viewBox="%s %s %s %s">
%s
</svg>
"""


def contour_sequences_to_svg(contour_sequences, html_width=300):
    "Contours as SVG paths (same output as triangulated.py:16-50)."
    lo = hi = None
    paths = []
    for (closed, sequence) in contour_sequences:
        parts = []
        for n, point in enumerate(sequence):
            parts.append(("L" if n else "M") + "%4.2f %4.2f" % tuple(point))
            point = np.array(point)
            lo = point if lo is None else np.min([point, lo], axis=0)
            hi = point if hi is None else np.max([point, hi], axis=0)
        if closed:
            parts.append("Z")
        paths.append(" ".join(parts))
    width_str = "%4.2f" % (0.01 * np.max(hi - lo))
    elements = ['<path stroke-width="%s" stroke="black" fill="none" d="%s" />' % (width_str, p) for p in paths]
    width, height = (hi - lo)
    scale = html_width * (1.0 / width)
    return SVG_TEMPLATE % (height * scale, html_width, lo[0], lo[1], width, height, "\n".join(elements))


def adjacent_pairs(low_pair, high_pair):
    "The 4 candidate keys adjacent to (low, high): they share its low or its high end (triangulated.py:66-77)."
    low = np.array(low_pair, dtype=int)
    high = np.array(high_pair, dtype=int)
    n = len(adjacent_offsets)
    low_index = adjacent_offsets.index(tuple(low - high))
    high_index = adjacent_offsets.index(tuple(high - low))
    for (hshift, lshift) in [(-1, 0), (1, 0), (0, -1), (0, 1)]:
        yield (tuple(int(x) for x in high + adjacency_array[(low_index + lshift) % n]),
               tuple(int(x) for x in low + adjacency_array[(high_index + hshift) % n]))


def chain_segments(keys2, pos2):
    """Polylines from a segment soup.  keys2 [S,2] uint64 end keys, pos2 [S,2,2] end points.
    Returns [(closed, points[k,2])].  Open polylines start at their smaller end key; closed ones at their smallest
    key, towards the smaller neighbour.  Consecutive np.allclose points are dropped (triangulated.py:269)."""
    if len(keys2) == 0:
        return []
    flat = keys2.reshape(-1)
    uk, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    pts = pos2.reshape(-1, 2)[first]
    ends = inv.reshape(-1, 2)
    # unique undirected segments
    e = np.unique(np.sort(ends, axis=1), axis=0)
    e = e[e[:, 0] != e[:, 1]]
    nk = len(uk)
    deg = np.bincount(e.reshape(-1), minlength=nk)
    # adjacency in CSR form
    both = np.concatenate([e, e[:, ::-1]])
    order = np.lexsort((both[:, 1], both[:, 0]))
    both = both[order]
    start = np.searchsorted(both[:, 0], np.arange(nk + 1))
    nbr = both[:, 1]
    visited = np.zeros(nk, dtype=bool)
    out = []

    def walk(s):
        chain = [s]
        visited[s] = True
        cur = s
        while True:
            nxt = -1
            for x in nbr[start[cur]:start[cur + 1]]:
                if not visited[x]:
                    nxt = x
                    break
            if nxt < 0:
                return chain
            visited[nxt] = True
            chain.append(nxt)
            cur = nxt

    for s in list(np.nonzero(deg < 2)[0]) + list(np.nonzero(deg >= 2)[0]):
        if visited[s]:
            continue
        closed = deg[s] >= 2
        chain = walk(s)
        p = pts[chain]
        if len(p) > 1:
            keep = np.ones(len(p), dtype=bool)
            keep[1:] = ~np.all(np.abs(p[1:] - p[:-1]) <= 1e-8 + 1e-5 * np.abs(p[:-1]), axis=1)
            p = p[keep]
        if len(p) > 1 and np.allclose(p[0], p[-1]):
            closed = True
        out.append((bool(closed), p))
    return out


class Grid2DContour(object):

    def __init__(self, horizontal_n, vertical_m, function, value, segment_endpoints=None, callback=None,
                 origin=(0.0, 0.0), delta=(1.0, 1.0)):
        self.n = horizontal_n
        self.m = vertical_m
        self.corner = np.array([horizontal_n, vertical_m], dtype=int)
        self.f = function
        self.z = value
        self.end_points = segment_endpoints
        self.callback = callback
        self.contours = []
        self.origin = origin
        self.delta = delta
        self.segments = None

    def _field(self):
        if isinstance(self.f, np.ndarray):
            if tuple(self.f.shape) != (self.n, self.m):
                raise ValueError("sample array has shape %r, expected %r" % (self.f.shape, (self.n, self.m)))
            return self.f
        g = field2d.Function2DGrid(0, 0, self.n - 1, self.m - 1, 1, 1, self.f)
        return g.samples(0)

    def get_contour_sequences(self):
        eng = E.default_engine()
        eng.mt2d_run(self._field(), [float(self.z)], origin=self.origin, delta=self.delta, flags=E.GEOM_F64)
        self.segments = eng.mt2d_fetch()
        self.contours = [(closed, pts) for closed, pts in chain_segments(self.segments["keys"], self.segments["pos"])]
        if self.callback:
            self.callback(self)
        return self.contours


class ContourGrid(object):
    "Shared 2D front end (triangulated.py:79-118)."

    def __init__(self, function_grid, value, segment_endpoints=None, linear_interpolate=True):
        self.linear_interpolate = linear_interpolate
        self.grid = function_grid
        self.value = value
        self.segment_endpoints = segment_endpoints
        self.contour_maker = self.get_contour_maker(None)
        self.grid_values = None


class DxDy2DContourGrid(ContourGrid):

    def get_contour_maker(self, grid_endpoints):
        assert self.linear_interpolate, "non-linear interpolation not implemented yet for 2d"
        grid = self.grid
        (n, m) = (int(x) for x in grid.grid_dimensions)
        return Grid2DContour(n, m, grid.samples(0), self.value, grid_endpoints)

    def get_contour_sequences(self):
        self.grid_contours = self.contour_maker.get_contour_sequences()
        self.contours = [self.from_grid_contour(c) for c in self.grid_contours]
        return self.contours

    def from_grid_contour(self, contour):
        (closed, grid_points) = contour
        return (closed, [self.grid.from_grid_coordinates(p) for p in grid_points])


class DxDy2DContour(DxDy2DContourGrid):

    def __init__(self, xmin, ymin, xmax, ymax, dx, dy, function, value, segment_endpoints=None):
        function_grid = field2d.Function2DGrid(xmin, ymin, xmax, ymax, dx, dy, function)
        DxDy2DContourGrid.__init__(self, function_grid, value, segment_endpoints)
