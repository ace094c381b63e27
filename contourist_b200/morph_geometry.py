"""Morphing triangles container + wire format -- drop-in for contourist/morph_geometry.py (MorphTriangles).

`to_json` restates morph_geometry.py:91-128 with numpy (same keys, same integer quantisation
int((p - min) / scale), scale = max(range, epsilon) / maxint) and is what misc/morph_triangles.js:3-52 loads.
"""
import numpy as np

from . import surface_geometry, wire


class MorphTriangles(object):

    def __init__(self, points4d, segment_point_indices, triangle_segment_indices):
        self.points4d = points4d = np.array(points4d, dtype=float).reshape(-1, 4)
        t = points4d[:, -1]
        self.max_value = t.max() if len(t) else 0.0
        self.min_value = t.min() if len(t) else 0.0
        seg = np.array(segment_point_indices, dtype=np.int64).reshape(-1, 2)
        if len(seg):
            swap = t[seg[:, 0]] > t[seg[:, 1]]                       # morph_geometry.py:13-17: low t first
            seg = np.where(swap[:, None], seg[:, ::-1], seg)
        self.segment_point_indices = seg
        self.triangle_segment_indices = np.array(triangle_segment_indices, dtype=np.int64).reshape(-1, 3)
        self.triangle_max_t = None
        self.triangle_min_t = None

    def from_grid_coordinates(self, grid):
        return MorphTriangles(grid.from_grid_coordinates(self.points4d), self.segment_point_indices,
                              self.triangle_segment_indices)

    def compute_triangle_stats(self):
        "Applicable t range of each triangle (morph_geometry.py:69-89)."
        t = self.points4d[:, -1]
        seg, tri = self.segment_point_indices, self.triangle_segment_indices
        lo = t[seg[:, 0]][tri]
        hi = t[seg[:, 1]][tri]
        self.triangle_min_t = np.maximum(lo.max(axis=1), self.min_value)
        self.triangle_max_t = np.minimum(hi.min(axis=1), self.max_value)

    def orient_triangles(self):
        """Right-hand-rule orientation of every triangle at its own mid-life, outward per the reference's max-x
        rule applied to the segment midpoints (morph_geometry.py:49-59).  Time compatibility
        (morph_geometry.py:61-67) restricts which triangles are considered connected: two triangles sharing an
        edge are joined only if their t ranges overlap."""
        self.compute_triangle_stats()
        p, seg, tri = self.points4d, self.segment_point_indices, self.triangle_segment_indices
        if len(tri) == 0:
            return
        mid = 0.5 * (p[seg[:, 0], :3] + p[seg[:, 1], :3])
        geometry = surface_geometry.SurfaceGeometry(mid, tri)
        tmin, tmax = self.triangle_min_t, self.triangle_max_t

        def compatible(k1, k2):
            return np.maximum(tmin[k1], tmin[k2]) < np.minimum(tmax[k1], tmax[k2])
        self.triangle_segment_indices = np.array(geometry.orient_triangles(link_filter=compatible),
                                                 dtype=np.int64).reshape(-1, 3)

    def to_json(self, min_value=None, max_value=None, maxint=999999, epsilon=1e-4):
        points = self.points4d
        min_value = self.min_value if min_value is None else max(min_value, self.min_value)
        max_value = self.max_value if max_value is None else min(max_value, self.max_value)
        L = ["{\n", '"description": "Ordered 4d morphing triangles.",\n',
             '"max_value": %s,\n' % (max_value,), '"min_value": %s,\n' % (min_value,),
             '"counts": [%s, %s, %s],\n' % (len(points), len(self.segment_point_indices), len(self.triangle_segment_indices))]
        maxima = points.max(axis=0)
        minima = points.min(axis=0)
        diff = np.maximum(maxima - minima, epsilon)
        L.append('"shift": [%s, %s, %s, %s],\n' % tuple(minima))
        scale = diff / maxint
        L.append('"scale": [%s, %s, %s, %s],\n' % tuple(scale))
        invscale = (1.0 / scale).reshape((1, 4))
        positions = ((points - minima.reshape(1, 4)) * invscale).astype(int)
        L.append('"positions": %s,\n' % (flatten_json_list(positions),))
        L.append('"segments": %s,\n' % (flatten_json_list(self.segment_point_indices),))
        L.append('"triangles": %s\n' % (flatten_json_list(self.triangle_segment_indices),))
        L.append("}")
        return "".join(L)


def flatten_json_list(sequence, fmt=str):
    "[a,b,..,\\nc,d,..]: rows joined by ',\\n', values by ',' (morph_geometry.py:127-128)."
    arr = np.asarray(sequence)
    if arr.ndim == 2 and arr.size and np.issubdtype(arr.dtype, np.integer) and fmt is str:
        return wire.format_rows(arr)                                  # host threads of the library, same bytes
    return "[%s]" % (",\n".join(",".join(fmt(y) for y in x) for x in sequence),)


class MorphGeometry(object):
    """One interval of the legacy "sequence of morphing triangularizations" (morph_geometry.py:130-330): the triangles in
    which the tetrahedra that live through [min_value, max_value] cut the time slice, with the positions of their
    corners at both ends of the interval.  Built by morph_sequence(); `json_data()` / `to_json()` write what
    misc/morph_sequence.js:3-20 loads."""

    def __init__(self, min_value, max_value, start_positions, end_positions, triangles):
        self.min_value, self.max_value = min_value, max_value
        self.start_positions = start_positions           # [n, 3] float
        self.end_positions = end_positions
        self.triangles = triangles                       # [m, 3] ints into the positions, oriented outwards

    def json_data(self, integral=True, lists_only=True, epsilon=1e-5, maxint=9999):
        "morph_geometry.py:277-307"
        D = {"description": "Morphing triangularization."}
        start, end = self.start_positions, self.end_positions
        if integral:
            positions = np.vstack([start, end])
            minima = positions.min(axis=0)
            diff = np.maximum(positions.max(axis=0) - minima, epsilon)
            D["shift"] = [float(x) for x in minima]
            scale = diff / maxint
            D["scale"] = [float(x) for x in scale]
            invscale = (1.0 / scale).reshape((1, 3))
            start = ((start - minima) * invscale).astype(int)
            end = ((end - minima) * invscale).astype(int)
        D["start_positions"], D["end_positions"], D["triangles"] = start, end, self.triangles
        if lists_only:
            for slot in ("start_positions", "end_positions", "triangles"):
                D[slot] = [list(x) for x in D[slot]]
        D["min_value"], D["max_value"] = self.min_value, self.max_value
        return D

    def to_json(self):
        "morph_geometry.py:309-329, same bytes for the same arrays"
        D = self.json_data(integral=True, lists_only=False)
        L = ["{", '"description": "Morphing triangularization.",\n', '"max_value": %s,\n' % (self.max_value,),
             '"min_value": %s,\n' % (self.min_value,), '"scale": %s,\n' % (D["scale"],), '"shift": %s' % (D["shift"],)]
        for slot in ("start_positions", "end_positions", "triangles"):
            rows = np.asarray(D[slot]).reshape(-1, 3)
            text = wire.format_rows(rows, row_sep=",") if len(rows) else "[]"
            L.append(',\n"%s": %s' % (slot, text))
        L.append("}\n")
        return "".join(L)


def morph_sequence(points4d, tetrahedra, epsilon=1e-5):
    """pentatopes.py:370-413 iterate_morph_geometry + morph_geometry.py:130-275, vectorised per interval.

    The reference walks the sorted vertex times; between two consecutive distinct times [u, w] it slices every active
    tetrahedron at the midpoint (add_tetrahedron), orients the triangles of that slice outwards and interpolates their
    corners to u and to w.  A tetrahedron is active over [u, w] iff some vertex has t <= u and some vertex t >= w; its
    cut edges are the (low, high) vertex pairs; one low or one high vertex gives one triangle, two and two give the
    quad (p1, p2, x) + (p1, p2, y) with p1 the first cut edge in the order (a,b),(a,c),(a,d),(b,c),(b,d),(c,d) of the
    sorted vertex ids and p2 the cut edge disjoint from it (morph_geometry.py:175-186).  Yields MorphGeometry objects."""
    P = np.asarray(points4d, dtype=float).reshape(-1, 4)
    T = np.sort(np.asarray(tetrahedra, dtype=np.int64).reshape(-1, 4), axis=1)
    if len(T) == 0:
        return
    t = P[:, 3]
    # the distinct vertex times, as the reference's np.allclose scan over the sorted values finds them
    values = np.sort(t)
    keep = [0]
    for q in range(1, len(values)):
        if not np.allclose(values[keep[-1]], values[q]):
            keep.append(q)
    u = values[keep]
    tt = t[T]
    PAIRS = np.array([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)])
    for lo_t, hi_t in zip(u[:-1], u[1:]):
        if np.allclose(lo_t, hi_t):
            continue
        act = (tt.min(axis=1) <= lo_t) & (tt.max(axis=1) >= hi_t)
        if not act.any():
            continue
        A = T[act]
        low = t[A] <= lo_t                                             # [n, 4]
        cut = low[:, PAIRS[:, 0]] != low[:, PAIRS[:, 1]]               # [n, 6] cut edges, in the reference's pair order
        ncut = cut.sum(axis=1)
        tri_pairs = []                                                 # [m, 3, 2] vertex-id pairs (i < j)
        pa, pb = A[:, PAIRS[:, 0]], A[:, PAIRS[:, 1]]                  # [n, 6]
        three = ncut == 3
        if three.any():
            idx = np.nonzero(cut[three])[1].reshape(-1, 3)
            rows = np.nonzero(three)[0][:, None]
            tri_pairs.append(np.stack([pa[rows, idx], pb[rows, idx]], axis=2))
        four = ncut == 4
        if four.any():
            idx = np.nonzero(cut[four])[1].reshape(-1, 4)              # pair numbers of the 4 cut edges, ascending
            rows = np.nonzero(four)[0]
            first = idx[:, 0]
            # the cut edge disjoint from the first: pairs k and 5 - k are the disjoint ones in this numbering
            partner = 5 - first
            others = np.array([[x for x in r if x != f and x != p] for r, f, p in zip(idx.tolist(), first.tolist(), partner.tolist())])
            for col in (0, 1):
                sel = np.stack([first, partner, others[:, col]], axis=1)
                tri_pairs.append(np.stack([pa[rows[:, None], sel], pb[rows[:, None], sel]], axis=2))
        if not tri_pairs:
            continue
        tp = np.concatenate(tri_pairs)                                 # [m, 3, 2]
        nv = len(P) + 1
        code = tp[:, :, 0] * nv + tp[:, :, 1]
        ucode, inv = np.unique(code.reshape(-1), return_inverse=True)
        tris = np.unique(np.sort(inv.reshape(-1, 3), axis=1), axis=0)  # a set of frozensets in the reference
        i1, i2 = ucode // nv, ucode % nv
        swap = t[i1] > t[i2]                                           # interpolate_pair_3d: vertex1 = the earlier one
        v1, v2 = np.where(swap, i2, i1), np.where(swap, i1, i2)

        def at(value):
            t1, t2 = t[v1], t[v2]
            val = np.full(len(v1), float(value))
            outside = (val + epsilon < t1) | (val - epsilon > t2)      # force=True: snap to the nearer end
            val = np.where(outside, np.where(np.abs(val - t1) < np.abs(val - t2), t1, t2), val)
            diff = t2 - t1
            ratio = np.where(diff > epsilon, (val - t1) / np.where(diff > epsilon, diff, 1.0), 0.0)
            return P[v1, :3] + ratio[:, None] * (P[v2, :3] - P[v1, :3])
        mid = at(0.5 * (lo_t + hi_t))
        geometry = surface_geometry.SurfaceGeometry(mid, tris)
        oriented = np.array(geometry.orient_triangles(), dtype=np.int64).reshape(-1, 3)
        yield MorphGeometry(lo_t, hi_t, at(lo_t), at(hi_t), oriented)


def morph_sequence_json(morphs):
    "pentatopes.py:430-444: the whole sequence as one JSON text"
    morphs = list(morphs)
    L = ["{\n", '"description": "Sequence of morphing triangularizations.",\n',
         '"max_value": %s,\n' % (morphs[-1].max_value,), '"min_value": %s,\n' % (morphs[0].min_value,),
         '"number_of_morphs": %s,\n' % (len(morphs),),
         '"morph_descriptions": [\n%s]\n' % (",\n".join(m.to_json() for m in morphs),), "}"]
    return "".join(L)
