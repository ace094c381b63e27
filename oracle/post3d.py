"""ORACLE (test infrastructure, not product code) -- the reference's 3D mesh post-processing, made deterministic.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

What the reference does after enumerating the raw triangles (tetrahedral.py:541-552, default arguments):
  tetrahedral.py:190-215   quantize_interpolations(divisions=10000)   -> quantize
  tetrahedral.py:353-375   remove_tiny_simplices(epsilon=1e-4)        -> tiny
  surface_geometry.py:14-50  clean_triangles                          -> clean
  (then surface_geometry.py:52-140 orient_triangles: oracle/mt3d.py orient)

All three depend on CPython dict / set iteration order in the reference (which interpolation of a quantum survives,
which point a collapsed simplex merges to, which of two coincident vertices keeps its index).  This restatement fixes
the order ("what the reference would do if every dict / set were sorted by the engine's vertex rank"):

  * vertices are ranked by `rank` (default: the engine's numbering -- owner word of 32 samples, edge direction, k --
    computed from the edge keys by engine_rank); wherever the reference keeps "whichever came last / first", the
    vertex of SMALLEST rank is kept;
  * sequential passes whose later iterations see the edits of earlier ones (tiny's position overwrite, clean's
    vertex_map) become: decide every simplex from the positions before the pass, then merge connected clusters.

Parity status: PINNED modulo that order.  tests/test_oracle_post3d.py runs postprocess() on the raw meshes the
unmodified reference produced (tests/golden/mt3d_*.npz, c1_sphere65.npz: key_pos / tris) and compares with the final
meshes the reference produced from them (final_points / final_tris): identical counts, identical triangles as sets of
positions except where a quantum holds several vertices (the survivor differs, positions then differ by less than
one quantum); the test states the residue per fixture.
"""
import numpy as np


def engine_rank(keys, shape):
    """Rank of every vertex in the engine's deterministic numbering: vertices ordered by (owner word = 32 consecutive
    samples along k, edge direction d, k), include/contourist_b200.h "vertex ids"."""
    n0, n1, n2 = (int(s) for s in shape)
    keys = np.asarray(keys, dtype=np.uint64)
    lin = (keys >> np.uint64(3)).astype(np.int64)
    d = (keys & np.uint64(7)).astype(np.int64)
    k = lin % n2
    row = lin // n2
    W = (n2 + 31) // 32
    word = row * W + k // 32
    order = np.lexsort((k, d, word))
    rank = np.empty(len(keys), dtype=np.int64)
    rank[order] = np.arange(len(keys), dtype=np.int64)
    return rank


def _roots(n, pairs_a, pairs_b, rank):
    """Union-find over vertex ids; the root of a cluster is its vertex of smallest rank."""
    parent = np.arange(n, dtype=np.int64)
    a = np.asarray(pairs_a, dtype=np.int64)
    b = np.asarray(pairs_b, dtype=np.int64)
    while len(a):
        ra, rb = parent[a], parent[b]
        ne = ra != rb
        if not ne.any():
            break
        ra, rb = ra[ne], rb[ne]
        lo = np.where(rank[ra] < rank[rb], ra, rb)
        hi = np.where(rank[ra] < rank[rb], rb, ra)
        # several unions may target one `hi`: keep the best (smallest rank) parent, iterate to the fixed point
        order = np.argsort(rank[lo], kind="stable")[::-1]
        parent[hi[order]] = lo[order]
        while True:
            nxt = parent[parent]
            if np.array_equal(nxt, parent):
                break
            parent = nxt
    return parent


def quantize(pos, tris, corner, rank, divisions=10000):
    """tetrahedral.py:190-215.  Returns (rep[V] representative vertex of every vertex, kept triangle mask, mapped tris).
    Positions are NOT changed (the reference stores the merged table in self.interpolated, which nothing reads)."""
    corner = np.asarray(corner, dtype=np.int64)
    expander = ((divisions * 1.0) / corner).astype(np.int64)               # :192
    q = (np.asarray(pos, dtype=np.float64) * expander).astype(np.int64)    # :196 (truncation towards zero)
    _, inv = np.unique(q, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    best = np.full(inv.max() + 1 if len(inv) else 0, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(best, inv, rank)
    by_rank = np.empty(len(rank), dtype=np.int64)
    by_rank[rank] = np.arange(len(rank), dtype=np.int64)
    rep = by_rank[best[inv]] if len(inv) else np.zeros(0, np.int64)
    t = rep[np.asarray(tris, dtype=np.int64)]
    distinct = (t[:, 0] != t[:, 1]) & (t[:, 0] != t[:, 2]) & (t[:, 1] != t[:, 2])          # :208
    # simplex_sets is a set of frozensets: equal triples collapse (:209); the first in triangle order stays
    s = np.sort(t, axis=1)
    keep = distinct.copy()
    idx = np.nonzero(distinct)[0]
    if len(idx):
        _, first = np.unique(s[idx], axis=0, return_index=True)
        keep[:] = False
        keep[idx[first]] = True
    return rep, keep, t


def tiny(pos, tris, corner, rank, epsilon=1e-4):
    """tetrahedral.py:353-375 on triangles [T,3] of vertex ids.  Returns (new positions, kept mask)."""
    pos = np.asarray(pos, dtype=np.float64)
    invcorner = 1.0 / np.asarray(corner, dtype=np.float64)
    P = pos[tris]                                                           # [T,3,3]
    delta = (P.max(axis=1) - P.min(axis=1)) * invcorner
    is_tiny = delta.max(axis=1) < epsilon if len(tris) else np.zeros(0, bool)
    tt = tris[is_tiny]
    parent = _roots(len(pos), np.concatenate([tt[:, 0], tt[:, 0]]), np.concatenate([tt[:, 1], tt[:, 2]]), rank)
    return pos[parent], ~is_tiny


def clean(pos, tris, rank):
    """surface_geometry.py:14-50.  Returns (kept mask over tris, merged vertex id per vertex)."""
    pos = np.asarray(pos, dtype=np.float64)
    A, B, C = pos[tris[:, 0]], pos[tris[:, 1]], pos[tris[:, 2]]
    cross = np.cross(A - C, B - C) if len(tris) else np.zeros((0, 3))
    flat = np.all(np.abs(cross) <= 1e-8, axis=1)                            # np.allclose(cross, 0)
    pa, pb = [], []
    for a, b in ((0, 1), (0, 2), (1, 2)):
        i, j = tris[flat, a], tris[flat, b]
        same = np.all(np.abs(pos[i] - pos[j]) <= 1e-8 + 1e-5 * np.abs(pos[j]), axis=1)     # np.allclose(p_i, p_j)
        pa.append(i[same])
        pb.append(j[same])
    parent = _roots(len(pos), np.concatenate(pa), np.concatenate(pb), rank)
    t = parent[tris]
    keep = ~flat & (t[:, 0] != t[:, 1]) & (t[:, 0] != t[:, 2]) & (t[:, 1] != t[:, 2])
    return keep, parent


def postprocess(pos, tris, corner, rank, divisions=10000, epsilon=1e-4, mins=(0.0, 0.0, 0.0), delta=(1.0, 1.0, 1.0)):
    """quantize -> tiny -> clean on a raw indexed mesh in GRID coordinates; vertices renumbered in rank order over the
    vertices the kept triangles use; positions mapped to world coordinates at the end (grid_field.py:89-93, as
    tetrahedral.py:86-90 does).  Returns dict(points [P,3], tris [Q,3], src [P] raw vertex id of every final point,
    tri_src [Q] raw triangle index, counts of what each pass removed)."""
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    tris = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    rank = np.asarray(rank, dtype=np.int64)
    rep, keep_q, t = quantize(pos, tris, corner, rank, divisions)
    tri_idx = np.nonzero(keep_q)[0]
    t = t[keep_q]
    pos2, keep_t = tiny(pos, t, corner, rank, epsilon)
    n_tiny = int((~keep_t).sum())
    tri_idx, t = tri_idx[keep_t], t[keep_t]
    keep_c, parent = clean(pos2, t, rank)
    n_flat = int((~keep_c).sum())
    tri_idx, t = tri_idx[keep_c], parent[t[keep_c]]
    used = np.unique(t)
    used = used[np.argsort(rank[used], kind="stable")]
    newid = -np.ones(len(pos), dtype=np.int64)
    newid[used] = np.arange(len(used))
    points = pos2[used] * np.asarray(delta, dtype=np.float64) + np.asarray(mins, dtype=np.float64)
    return dict(points=points, tris=newid[t], src=used, tri_src=tri_idx,
                n_quantized=int((~keep_q).sum()), n_tiny=n_tiny, n_flat=n_flat)
