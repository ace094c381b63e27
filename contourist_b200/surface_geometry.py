"""Triangle mesh container -- drop-in for contourist/surface_geometry.py (SurfaceGeometry).

The engine already delivers an indexed mesh whose triangles are wound consistently (normal towards the high
side of the field), so `orient_triangles` only has to decide one flip per connected component to reproduce the
reference's convention (surface_geometry.py:79-103: at the component's max-x vertex the triangle with the
largest |cross.x| gets cross.x > 0).  Components are found with a vectorised union-find over triangle edges.
`clean_triangles` restates surface_geometry.py:14-50 with numpy (zero-area triangles dropped, coincident
vertices of those triangles merged, vertices renumbered in first-use order of the sorted triangle list).
"""
import numpy as np


def _components(n_verts, tris):
    """Connected components of the vertex graph induced by triangles (pointer-jumping union-find, vectorised)."""
    parent = np.arange(n_verts, dtype=np.int64)
    if len(tris) == 0:
        return parent
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [0, 2]]])
    while True:
        ra, rb = parent[e[:, 0]], parent[e[:, 1]]
        lo, hi = np.minimum(ra, rb), np.maximum(ra, rb)
        changed = lo != hi
        if not changed.any():
            break
        np.minimum.at(parent, hi[changed], lo[changed])
        while True:                      # compress
            nxt = parent[parent]
            if np.array_equal(nxt, parent):
                break
            parent = nxt
    return parent


class SurfaceGeometry(object):

    def __init__(self, vertices, triangles):
        self.input_vertices = vertices
        self.input_triangles = triangles
        self.vertices = vertices
        self.triangles = triangles
        self.oriented_triangles = triangles
        self.vertex_map = tuple(range(len(vertices)))

    def _arrays(self, vertices, triangles):
        V = np.asarray(vertices, dtype=float).reshape(-1, 3)
        T = np.array([tuple(t) for t in triangles], dtype=np.int64).reshape(-1, 3) if not isinstance(triangles, np.ndarray) \
            else triangles.astype(np.int64).reshape(-1, 3)
        return V, T

    def clean_triangles(self):
        "Remove area 0 triangles and merge coincident vertices of those triangles (surface_geometry.py:14-50)."
        V, T = self._arrays(self.input_vertices, self.input_triangles)
        A, B, C = V[T[:, 0]], V[T[:, 1]], V[T[:, 2]]
        cross = np.cross(A - C, B - C)
        flat = np.all(np.abs(cross) <= 1e-8, axis=1)                   # np.allclose(cross, 0)
        # merge identical vertices of the dropped triangles (union-find on allclose pairs)
        parent = np.arange(len(V), dtype=np.int64)
        for a, b in ((0, 1), (0, 2), (1, 2)):
            i, j = T[flat, a], T[flat, b]
            same = np.all(np.abs(V[i] - V[j]) <= 1e-8 + 1e-5 * np.abs(V[j]), axis=1)
            for x, y in zip(i[same], j[same]):
                rx, ry = x, y
                while parent[rx] != rx:
                    rx = parent[rx]
                while parent[ry] != ry:
                    ry = parent[ry]
                if rx != ry:
                    parent[max(rx, ry)] = min(rx, ry)
        while True:
            nxt = parent[parent]
            if np.array_equal(nxt, parent):
                break
            parent = nxt
        keep = parent[T[~flat]]
        keep = keep[(keep[:, 0] != keep[:, 1]) & (keep[:, 0] != keep[:, 2]) & (keep[:, 1] != keep[:, 2])]
        used, inv = np.unique(keep, return_inverse=True)
        vmap = -np.ones(len(V), dtype=np.int64)
        vmap[used] = np.arange(len(used))
        self.vertices = V[used]
        self.triangles = inv.reshape(-1, 3).astype(np.int64)
        self.vertex_map = {int(i): int(vmap[parent[i]]) for i in range(len(V)) if vmap[parent[i]] >= 0}
        self.oriented_triangles = self.triangles
        return (self.vertices, self.triangles)

    def orient_triangles(self, compatible_triangle_test=None, link_filter=None):
        """Consistent winding per connected component + the reference's outward rule (surface_geometry.py:52-140):
        in every component the triangle with the largest |cross.x| at the max-x vertex gets cross.x > 0.
        Winding is propagated across shared edges by solving the parity constraints on a doubled graph
        (vectorised; the reference walks a DFS).  `link_filter(k1, k2) -> bool array` may veto links between triangle
        numbers (MorphTriangles uses it for time compatibility, morph_geometry.py:61-67); the reference's own form,
        `compatible_triangle_test(triangle1, triangle2) -> bool` on two triangles (tuples of vertex indices, as stored in
        self.triangles; surface_geometry.py:52-56,118), is honoured the same way, one call per linked pair."""
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        V, T = self._arrays(self.vertices, self.triangles)
        K = len(T)
        if K == 0:
            self.oriented_triangles = []
            return self.oriented_triangles
        nv = int(T.max()) + 1
        u = T.reshape(-1)
        v = T[:, [1, 2, 0]].reshape(-1)
        tri_of = np.repeat(np.arange(K), 3)
        key = np.minimum(u, v) * nv + np.maximum(u, v)
        order = np.argsort(key, kind="stable")
        ks, us, ts = key[order], u[order], tri_of[order]
        same_edge = ks[1:] == ks[:-1]
        k1, k2 = ts[:-1][same_edge], ts[1:][same_edge]
        rel = (us[:-1][same_edge] == us[1:][same_edge]).astype(np.int64)      # same direction -> one must flip
        if link_filter is None and compatible_triangle_test is not None:
            as_given = [tuple(int(x) for x in t) for t in T]

            def link_filter(a, b):
                return np.array([bool(compatible_triangle_test(as_given[i], as_given[j])) for i, j in zip(a, b)], dtype=bool)
        if link_filter is not None and len(k1):
            ok = link_filter(k1, k2)
            k1, k2, rel = k1[ok], k2[ok], rel[ok]
        rows = np.concatenate([2 * k1, 2 * k1 + 1])
        cols = np.concatenate([2 * k2 + rel, 2 * k2 + (1 - rel)])
        graph = coo_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(2 * K, 2 * K))
        _, lab2 = connected_components(graph, directed=False)
        base = coo_matrix((np.ones(len(k1), dtype=np.int8), (k1, k2)), shape=(K, K))
        ncomp, lab = connected_components(base, directed=False)
        # per base component: the doubled component holding (first triangle, unflipped)
        first = np.full(ncomp, K, dtype=np.int64)
        np.minimum.at(first, lab, np.arange(K))
        flip = (lab2[2 * np.arange(K)] != lab2[2 * first[lab]]).astype(bool)
        out = np.where(flip[:, None], T[:, ::-1], T)
        # outward rule per component
        cross_x = np.cross(V[out[:, 0]] - V[out[:, 1]], V[out[:, 0]] - V[out[:, 2]])[:, 0]
        vx = V[:, 0]
        tri_maxx = vx[out].max(axis=1)
        comp_maxx = np.full(ncomp, -np.inf)
        np.maximum.at(comp_maxx, lab, tri_maxx)
        at_top = tri_maxx == comp_maxx[lab]
        score = np.where(at_top, np.abs(cross_x), -1.0)
        best = np.full(ncomp, -1.0)
        np.maximum.at(best, lab, score)
        is_best = at_top & (score == best[lab])
        seed = np.full(ncomp, K, dtype=np.int64)
        np.minimum.at(seed, lab[is_best], np.nonzero(is_best)[0])
        comp_flip = cross_x[seed] < 0
        out = np.where(comp_flip[lab][:, None], out[:, ::-1], out)
        self.oriented_triangles = sorted(tuple(int(x) for x in row) for row in out)
        return self.oriented_triangles
