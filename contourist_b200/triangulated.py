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


def _graph(keys2, pos2):
    "vertices (unique end keys, in key order), their points, unique undirected edges, degrees"
    flat = keys2.reshape(-1)
    uk, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    pts = pos2.reshape(-1, 2)[first]
    ends = inv.reshape(-1, 2)
    e = np.unique(np.sort(ends, axis=1), axis=0)
    e = e[e[:, 0] != e[:, 1]]
    deg = np.bincount(e.reshape(-1), minlength=len(uk))
    return uk, pts, e, deg


def _finish(closed, p):
    "consecutive np.allclose points dropped (triangulated.py:269); a chain that returns to its start is closed"
    if len(p) > 1:
        keep = np.ones(len(p), dtype=bool)
        keep[1:] = ~np.all(np.abs(p[1:] - p[:-1]) <= 1e-8 + 1e-5 * np.abs(p[:-1]), axis=1)
        p = p[keep]
    if len(p) > 1 and np.allclose(p[0], p[-1]):
        closed = True
    return (bool(closed), p)


def _chain_walk(pts, e, deg):
    """The per-vertex walk (triangulated.py:236-293 with a fixed order): used when some key has more than two
    neighbours (a sample exactly on the level), where the reference's result is a matter of visiting order."""
    nk = len(deg)
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
        out.append(_finish(deg[s] >= 2, pts[walk(s)]))
    return out


def rank_darts(e, deg):
    """List ranking of the contours of a graph whose vertices have at most two neighbours (paths and cycles), without
    walking them (triangulated.py:221-305 does it a vertex at a time).  Every edge {u, v} becomes two darts u->v, v->u;
    the successor of a dart arriving at b is the dart that leaves b along b's OTHER edge, or -- at a path's end -- the
    dart back along the same edge.  A path of m edges is then one cycle of 2m darts (there and back), a closed contour
    of m edges two cycles of m darts (one per direction).  Pointer jumping (log2 of the longest contour rounds, each a
    few gathers over all darts) gives every dart the first dart of its cycle and its position behind it:
      first dart = the dart leaving the path's smaller end / the cycle's smallest vertex (vertex numbers are key ranks);
      of a closed contour's two directions the one that starts towards the smaller neighbour is kept;
      of a path's there-and-back tour the first half.
    Returns (tail, head, start dart of each dart's cycle, position, cycle length, keep mask, open mask)."""
    ne = len(e)
    nd = 2 * ne
    tail = np.empty(nd, dtype=np.int64)
    head = np.empty(nd, dtype=np.int64)
    tail[0::2], head[0::2] = e[:, 0], e[:, 1]
    tail[1::2], head[1::2] = e[:, 1], e[:, 0]
    # the (at most two) darts that leave each vertex
    order = np.argsort(tail, kind="stable")
    first = np.searchsorted(tail[order], np.arange(len(deg)))
    out0 = order[np.minimum(first, nd - 1)]
    out1 = np.where(deg >= 2, order[np.minimum(first + 1, nd - 1)], -1)
    d = np.arange(nd, dtype=np.int64)
    rev = d ^ 1
    o0, o1 = out0[head], out1[head]
    succ = np.where(o0 != rev, o0, o1)
    succ = np.where(succ < 0, rev, succ)                   # a path's end: turn round
    # first dart of every cycle: smallest (not leaving an end, tail vertex), carried with the dart's own number
    prio = (np.where(deg[tail] == 1, 0, 1).astype(np.int64) * (len(deg) + 1) + tail) * nd + d
    best, jump = prio.copy(), succ.copy()
    while True:
        nb = np.minimum(best, best[jump])
        jump = jump[jump]
        if np.array_equal(nb, best):
            break
        best = nb
    start = best % nd
    # position behind the first dart: cut every cycle in front of its first dart, rank the lists
    nxt = np.where(succ == start, nd, succ)
    dist = np.where(nxt == nd, 0, 1).astype(np.int64)
    nxt = np.append(nxt, nd)
    dist = np.append(dist, 0)
    while True:
        live = nxt[:nd] != nd
        if not live.any():
            break
        dist[:nd] += dist[nxt[:nd]]
        nxt[:nd] = nxt[nxt[:nd]]
    dist = dist[:nd]
    length = dist[start] + 1
    pos = length - 1 - dist
    twin = start[rev]
    is_open = twin == start
    keep = np.where(is_open, pos < length // 2, head[start] < head[twin])
    return tail, head, start, pos, length, keep, is_open


def chain_segments(keys2, pos2):
    """Polylines from a segment soup.  keys2 [S,2] uint64 end keys, pos2 [S,2,2] end points.
    Returns [(closed, points[k,2])], open polylines first.  Open polylines start at their smaller end key; closed ones
    at their smallest key, towards the smaller neighbour.  Consecutive np.allclose points are dropped
    (triangulated.py:269).  Vectorised list ranking (rank_darts); the per-vertex walk only where a key has more than
    two neighbours."""
    if len(keys2) == 0:
        return []
    uk, pts, e, deg = _graph(np.asarray(keys2), np.asarray(pos2))
    if len(e) == 0:
        return []
    if deg.max() > 2:
        return _chain_walk(pts, e, deg)
    tail, head, start, pos, length, keep, is_open = rank_darts(e, deg)
    # kept darts grouped by contour: open ones first (by their first vertex), then closed ones, each in position order
    kd = np.nonzero(keep)[0]
    closed_d = ~is_open[kd]
    grp = np.lexsort((pos[kd], tail[start[kd]], closed_d))
    kd = kd[grp]
    first_of = np.nonzero(pos[kd] == 0)[0]                 # one per contour, in output order
    bounds = np.append(first_of, len(kd))
    out = []
    for q in range(len(first_of)):
        dd = kd[bounds[q]:bounds[q + 1]]
        closed = not is_open[dd[0]]
        verts = np.concatenate([[tail[dd[0]]], head[dd[:-1]] if closed else head[dd]])
        out.append(_finish(closed, pts[verts]))
    return out


def seeded_segments(value_at, z, n1, seg_keys, end_points):
    """Which segments of a full scan the reference's tracker reaches from seed segments (triangulated.py:307-338
    find_initial_contour_pairs / expand_contour_pairs): every seed is bisected down to adjacent points (a pair
    "exists" iff f(low) <= z <= f(high)); the start pairs are the contour pairs with the low point as their low end or
    the high point as their high end; a pair brings in every pair that shares its low end (as low end) or its high
    end (as high end).  seg_keys [S,2] uint64 engine keys ((lin(min point)*4 + d) << 1 | lowmin), n1 = row length,
    value_at(point) -> f.  Returns a bool mask over the segments."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    seg_keys = np.asarray(seg_keys, dtype=np.uint64).reshape(-1, 2)
    keys, inv = np.unique(seg_keys.reshape(-1), return_inverse=True)
    if len(keys) == 0:
        assert len(end_points) == 0, "bad end points: the level has no contour"
        return np.zeros(0, dtype=bool)
    lowmin = (keys & np.uint64(1)).astype(np.int64)
    kd = keys >> np.uint64(1)
    d = (kd & np.uint64(3)).astype(np.int64)
    lin_p = (kd >> np.uint64(2)).astype(np.int64)
    lin_q = lin_p + ((d >> 1) & 1) * n1 + (d & 1)
    low_lin = np.where(lowmin == 1, lin_p, lin_q)
    high_lin = np.where(lowmin == 1, lin_q, lin_p)
    rows, cols = [], []
    for end in (low_lin, high_lin):                        # pairs sharing an end: consecutive after sorting by it
        o = np.argsort(end, kind="stable")
        same = end[o][1:] == end[o][:-1]
        rows.append(o[:-1][same])
        cols.append(o[1:][same])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    _, comp = connected_components(coo_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(len(keys), len(keys))),
                                   directed=False)

    def exists(lo, hi):
        return value_at(lo) <= z <= value_at(hi)

    wanted = set()
    for low_point, high_point in np.array(end_points, dtype=int).reshape(-1, 2, 2):
        if not exists(low_point, high_point):
            (low_point, high_point) = (high_point, low_point)
            assert exists(low_point, high_point), "bad end points " + repr((tuple(low_point), tuple(high_point)))
        while np.any(np.abs(low_point - high_point) > 1):
            mid_point = (low_point + high_point) // 2
            if exists(low_point, mid_point):
                high_point = mid_point
            else:
                assert exists(mid_point, high_point)
                low_point = mid_point
        start = (low_lin == low_point[0] * n1 + low_point[1]) | (high_lin == high_point[0] * n1 + high_point[1])
        assert start.any()
        wanted.update(comp[start].tolist())
    return np.isin(comp[inv.reshape(-1, 2)[:, 0]], sorted(wanted))


class Grid2DContour(object):
    """Grid-coordinate front end.  segment_endpoints=None: every contour of the level (full scan); a list of seed
    segments: only the contours the reference's tracker reaches from them (seeded_segments)."""

    full_scan = False                                  # True: ignore the seeds (they are known to be every crossing)

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

    def _value_at(self, point):
        "f at an integer point: the callable wherever it is defined (seed segments may end outside the grid), else the array"
        if not isinstance(self.f, np.ndarray):
            return float(self.f(*(int(x) for x in point)))
        i, j = int(point[0]), int(point[1])
        if not (0 <= i < self.n and 0 <= j < self.m):
            raise ValueError("seed point %r outside the %d x %d sample array" % ((i, j), self.n, self.m))
        return float(self.f[i, j])

    def get_contour_sequences(self):
        eng = E.default_engine()
        eng.mt2d_run(self._field(), [float(self.z)], origin=self.origin, delta=self.delta, flags=E.GEOM_F64)
        self.segments = eng.mt2d_fetch()
        if self.end_points is None or self.full_scan:
            polys = eng.mt2d_polylines()                   # chained on the device; None: a key on more than two segments
            if polys is not None:
                self.contours = polys[0]
                if self.callback:
                    self.callback(self)
                return self.contours
        if self.end_points is not None and not self.full_scan:
            keep = seeded_segments(self._value_at, float(self.z), self.m, self.segments["keys"], self.end_points)
            self.segments = {k: (v[keep] if isinstance(v, np.ndarray) and len(v) == len(keep) else v)
                             for k, v in self.segments.items()}
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
        grid_endpoints = None
        if segment_endpoints is not None:               # triangulated.py:92-103: world seeds -> grid seeds, else grid search
            grid_endpoints = []
            for (start_xy, end_xy) in segment_endpoints:
                assert len(start_xy) == len(end_xy) == 2
                grid_endpoint = self.to_grid_endpoint(np.array(start_xy, dtype=float), np.array(end_xy, dtype=float))
                if grid_endpoint is not None:
                    grid_endpoints.append(grid_endpoint)
            if len(grid_endpoints) < 1:
                grid_endpoints = None
        self.contour_maker = self.get_contour_maker(grid_endpoints)
        self.grid_values = None

    def to_grid_endpoint(self, start_xy, end_xy):
        "triangulated.py:109-118: the first pair of grid vertices around the two points that straddles the value."
        grid, value = self.grid, self.value
        for start_grid in grid.surrounding_vertices(start_xy):
            for end_grid in grid.surrounding_vertices(end_xy):
                if not np.all(start_grid == end_grid):
                    if (grid.grid_function(*start_grid) - value) * (grid.grid_function(*end_grid) - value) <= 0:
                        return (start_grid, end_grid)
        return None


class DxDy2DContourGrid(ContourGrid):

    def get_contour_maker(self, grid_endpoints):
        assert self.linear_interpolate, "non-linear interpolation not implemented yet for 2d"
        grid = self.grid
        (n, m) = (int(x) for x in grid.grid_dimensions)
        return Grid2DContour(n, m, grid.samples(0), self.value, grid_endpoints)    # None: grid search = full scan

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
