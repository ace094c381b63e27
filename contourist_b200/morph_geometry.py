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
