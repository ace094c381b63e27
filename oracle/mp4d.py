"""ORACLE (test infrastructure, not product code) -- 4D marching pentatopes + morph triangles.

numpy restatement of the reference's 4D path as a FULL SCAN.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Parity status: PINNED for the raw extraction (hypervoxels, case codes, keys, positions, tetrahedra modulo
the reference's set-order split of 2-3 prisms) and for bin_times / drop_instant_tetrahedra / the per-tet
slicing, against golden vectors made by running the unmodified reference (tests/golden/make_golden.py ->
mp4d_*.npz).  The reference has no tests or known answers for this path.

Reference lines restated (under /root/reference/contourist/):
  pentatopes.py:15-39     PENTATOPES (24 Kuhn pentatopes), HYPERCUBE
  tetrahedral.py:383-394  border_voxel with the 16-corner box
  pentatopes.py:223-291   enumerate_pentatope_tetrahedra: 1-vs-4 -> {ab,ac,ad,ae};
                          2-vs-3 -> {ac,be,ad,bd}, {ac,be,ad,ae}, {ac,be,bd,bc}
  tetrahedral.py:471-487  contour_pair_interpolation (shared with 3D)
  pentatopes.py:162-169   bin_times(100)
  pentatopes.py:171-189   drop_instant_tetrahedra(1e-7)
  tetrahedral.py:353-375  remove_tiny_simplices(1e-3)
  morph_geometry.py:145-237  triangulate_tetrahedron_at_midpoints / add_tetrahedron / interpolate_pair_3d
  pentatopes.py:314-368   collect_morph_triangles
  morph_geometry.py:5-22,91-128  MorphTriangles, to_json

Conventions shared with the CUDA engine: field[i,j,k,l], corner number c = di*8+dj*4+dk*2+dl, edge key =
lin(min endpoint)*16 + d, d in 1..15; per-pentatope case code = 5-bit low mask (bit b <-> vertex b of
PENTATOPES[p]) | 32 if skipped by allclose; [a, b] / [c, d, e] are taken in sorted (lexicographic) order.
"""
import itertools

import numpy as np

RTOL = 1e-5
ATOL = 1e-8


def _pentatopes():
    out = []
    for perm in itertools.permutations(range(4)):
        v = [0, 0, 0, 0]
        verts = [0]
        for idx in perm:
            v[idx] = 1
            verts.append(v[0] * 8 + v[1] * 4 + v[2] * 2 + v[3])
        out.append(verts)
    return np.array(out, dtype=np.int64)


PENTS = _pentatopes()                    # pentatopes.py:15-26 order
CORNERS = np.array([[(c >> 3) & 1, (c >> 2) & 1, (c >> 1) & 1, c & 1] for c in range(16)], dtype=np.int64)


def _corner_views(field):
    n = field.shape
    out = []
    for d in CORNERS:
        out.append(field[d[0]:n[0] - 1 + d[0], d[1]:n[1] - 1 + d[1], d[2]:n[2] - 1 + d[2], d[3]:n[3] - 1 + d[3]])
    return out


def pent_cases(field, value):
    """uint8 [24, cells...]: 5-bit low mask | 32 if np.allclose(values, value) (pentatopes.py:229-238)."""
    value = float(value)
    cs = _corner_views(field)
    low = [c.astype(np.float64) < value for c in cs]
    nv = [np.abs(c.astype(np.float64) - value) <= ATOL + RTOL * abs(value) for c in cs]
    out = np.zeros((24,) + cs[0].shape, dtype=np.uint8)
    for p, pent in enumerate(PENTS):
        alln = np.ones(cs[0].shape, dtype=bool)
        for b, corner in enumerate(pent):
            out[p] |= (low[corner].astype(np.uint8) << b)
            alln &= nv[corner]
        out[p] |= alln.astype(np.uint8) << 5
    return out


def active_cells(field, value):
    """tetrahedral.py:383-394 with the 16-corner box."""
    value = float(value)
    cs = _corner_views(field)
    mn = cs[0].astype(np.float64)
    mx = mn.copy()
    fv0 = cs[0].astype(np.float64)
    allnear = np.abs(value - fv0) <= ATOL + RTOL * np.abs(fv0)
    for c in cs[1:]:
        c64 = c.astype(np.float64)
        mn = np.minimum(mn, c64)
        mx = np.maximum(mx, c64)
        allnear &= np.abs(value - c64) <= ATOL + RTOL * np.abs(c64)
    return (~allnear) & (mn <= value) & (mx >= value)


def extract(field, value, geom_dtype=np.float64):
    """Raw full-scan extraction (before bin_times).  Returns dict:
      cells int64 [Nc] linear hypervoxel index (over the cell grid), sorted; codes uint8 [Nc,24]
      keys uint64 [V] sorted unique; lowmin uint8 [V]; pos [V,4] grid coordinates
      tet_keys uint64 [T,4] in (cell, pentatope, split) order; tets int32 [T,4]; tet_cell int64 [T]; tet_pent uint8 [T]
    """
    field = np.asarray(field)
    value = float(value)
    n = field.shape
    act = active_cells(field, value)
    codes = pent_cases(field, value)
    m = codes & 31
    emits = (m != 0) & (m != 31) & ((codes & 32) == 0)
    sel = act & emits.any(axis=0)
    idx = np.nonzero(sel)
    cn = [s - 1 for s in n]
    cell_lin = ((idx[0] * cn[1] + idx[1]) * cn[2] + idx[2]) * cn[3] + idx[3]
    lin0 = ((idx[0] * n[1] + idx[1]) * n[2] + idx[2]) * n[3] + idx[3]
    corner_lin = [lin0 + ((int(d[0]) * n[1] + int(d[1])) * n[2] + int(d[2])) * n[3] + int(d[3]) for d in CORNERS]
    codes_sel = codes[:, idx[0], idx[1], idx[2], idx[3]].T.copy()      # [Nc,24]

    def key_of(ca, cb, rows):
        lo_c, hi_c = (ca, cb) if ca < cb else (cb, ca)
        return (corner_lin[lo_c][rows].astype(np.uint64) << np.uint64(4)) | np.uint64(hi_c - lo_c)

    tk, tc, tp, to = [], [], [], []
    for p, pent in enumerate(PENTS):
        c6 = codes_sel[:, p]
        em = ((c6 & 31) != 0) & ((c6 & 31) != 31) & ((c6 & 32) == 0)
        for mask in range(1, 31):
            rows = np.nonzero(em & ((c6 & 31) == mask))[0]
            if rows.size == 0:
                continue
            lows = sorted(int(pent[b]) for b in range(5) if (mask >> b) & 1)
            highs = sorted(int(pent[b]) for b in range(5) if not (mask >> b) & 1)
            least, most = lows, highs
            if len(least) > len(most):
                least, most = most, least
            if len(least) == 1:
                a = least[0]
                b_, c_, d_, e_ = most
                tl = [((a, b_), (a, c_), (a, d_), (a, e_))]
            else:
                a, b_ = least
                c_, d_, e_ = most
                ac, ad, ae, bc, bd, be = (a, c_), (a, d_), (a, e_), (b_, c_), (b_, d_), (b_, e_)
                tl = [(ac, be, ad, bd), (ac, be, ad, ae), (ac, be, bd, bc)]
            for s, tet in enumerate(tl):
                tk.append(np.stack([key_of(u, w, rows) for (u, w) in tet], axis=1))
                tc.append(cell_lin[rows])
                tp.append(np.full(rows.size, p, dtype=np.uint8))
                to.append(cell_lin[rows] * 72 + p * 3 + s)
    if tk:
        tk = np.concatenate(tk); tc = np.concatenate(tc); tp = np.concatenate(tp)
        o = np.argsort(np.concatenate(to), kind="stable")
        tk, tc, tp = tk[o], tc[o], tp[o]
    else:
        tk = np.zeros((0, 4), np.uint64); tc = np.zeros(0, np.int64); tp = np.zeros(0, np.uint8)
    keys = np.unique(tk)
    tets = np.searchsorted(keys, tk).astype(np.int32)
    lowmin, pos = interpolate(field, value, keys, geom_dtype)
    return dict(cells=cell_lin.astype(np.int64), codes=codes_sel, keys=keys, lowmin=lowmin, pos=pos,
                tet_keys=tk, tets=tets, tet_cell=tc, tet_pent=tp)


def key_points(keys, shape):
    n = shape
    lin = (keys >> np.uint64(4)).astype(np.int64)
    d = (keys & np.uint64(15)).astype(np.int64)
    l = lin % n[3]
    k = (lin // n[3]) % n[2]
    j = (lin // (n[3] * n[2])) % n[1]
    i = lin // (n[3] * n[2] * n[1])
    pmin = np.stack([i, j, k, l], axis=1)
    pmax = pmin + np.stack([(d >> 3) & 1, (d >> 2) & 1, (d >> 1) & 1, d & 1], axis=1)
    return pmin, pmax


def interpolate(field, value, keys, geom_dtype=np.float64):
    """tetrahedral.py:471-487 in 4D."""
    pmin, pmax = key_points(keys, field.shape)
    gd = np.dtype(geom_dtype)
    fmin = field[pmin[:, 0], pmin[:, 1], pmin[:, 2], pmin[:, 3]].astype(gd)
    fmax = field[pmax[:, 0], pmax[:, 1], pmax[:, 2], pmax[:, 3]].astype(gd)
    swap = fmin > fmax
    lowmin = (~swap).astype(np.uint8)
    flow = np.where(swap, fmax, fmin)
    fhigh = np.where(swap, fmin, fmax)
    plow = np.where(swap[:, None], pmax, pmin).astype(gd)
    phigh = np.where(swap[:, None], pmin, pmax).astype(gd)
    den = fhigh - flow
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (gd.type(value) - flow) / den
    ratio = np.where(np.abs(den.astype(np.float64)) <= ATOL, gd.type(0.5), ratio).astype(gd)
    return lowmin, (plow + ratio[:, None] * (phigh - plow)).astype(gd)


# ---------------------------------------------------------------------------------------------- post steps
def bin_times(pos, corner_t, nbins=100):
    """pentatopes.py:162-169: t snapped DOWN to multiples of corner_t/nbins (python int() truncation)."""
    pos = np.array(pos, dtype=np.float64)
    min_interval = corner_t * (1.0 / nbins)
    pos[:, -1] = np.trunc(pos[:, -1] / min_interval) * min_interval
    return pos


def drop_instant(pos, tets, epsilon=1e-7):
    """pentatopes.py:171-189: keep tets whose t extent is >= epsilon."""
    t = pos[:, -1][tets]
    return (t.max(axis=1) - t.min(axis=1)) >= epsilon


def tiny_mask(pos, tets, corner, epsilon=1e-3):
    """tetrahedral.py:353-375 predicate only (the reference also moves the tiny simplex's vertices to its first
    vertex, in set order -- second tier)."""
    p = pos[tets]                                    # [T,4,4]
    ext = (p.max(axis=1) - p.min(axis=1)) * (1.0 / np.asarray(corner, dtype=np.float64))
    return ext.max(axis=1) < epsilon


def slice_polygons(v4, tet, eps_gap=1e-4, eps_in=1e-5):
    """morph_geometry.py:145-153,202-227 for one tetrahedron: for every gap > eps_gap between consecutive sorted
    t values, the ordered list of tet edges (p, q), p < q, whose t range (widened by eps_in) contains the
    gap's midpoint.  3 edges = a triangle, 4 = a quad."""
    tet = sorted(int(x) for x in tet)
    if len(set(tet)) < 4:
        return []
    tv = sorted(v4[i][-1] for i in tet)
    out = []
    prev = None
    for cur in tv:
        if prev is not None and (cur - prev) > eps_gap:
            mid = 0.5 * (cur + prev)
            a, b, c, d = tet
            inter = []
            for (p, q) in ((a, b), (a, c), (a, d), (b, c), (b, d), (c, d)):
                v1, v2 = v4[p][-1], v4[q][-1]
                if v1 > v2:
                    v1, v2 = v2, v1
                if mid + eps_in < v1 or mid - eps_in > v2:
                    continue
                inter.append((p, q))
            out.append(inter)
        prev = cur
    return out


def slice_tet(v4, tet, eps_gap=1e-4, eps_in=1e-5):
    """morph_geometry.py:155-192 on top of slice_polygons.  Returns list of triangles, each a frozenset of
    3 vertex-id pairs (i<j)."""
    out = []
    for inter in slice_polygons(v4, tet, eps_gap, eps_in):
        if len(inter) == 3:
            out.append(frozenset(inter))
        elif len(inter) == 4:
            pair1 = inter[0]
            pair2 = None
            for pr in inter[1:]:
                if not (set(pr) & set(pair1)):
                    pair2 = pr
            assert pair2 is not None
            for pr in inter:
                if pr != pair1 and pr != pair2:
                    out.append(frozenset([pair1, pair2, pr]))
    return out


def morph_triangles(v4, tets, epsilon=1e-7):
    """pentatopes.py:314-368 (without the final orientation): returns (segments [M,2] low-t first,
    triangles [K,3] segment ids, sorted rows, rows sorted)."""
    v4 = np.asarray(v4, dtype=np.float64)
    tris = set()
    for tet in set(frozenset(int(x) for x in t) for t in tets):
        if len(tet) == 4:
            tris.update(slice_tet(v4, tet))
    tv = v4[:, -1]
    t_eps = epsilon * (tv.max() - tv.min()) if len(tv) else 0.0
    keep = []
    for tri in tris:
        if all(abs(v4[i][-1] - v4[j][-1]) > t_eps for (i, j) in tri):
            keep.append(tri)
    segs = sorted(set(p for tri in keep for p in tri))
    sid = {p: n for n, p in enumerate(segs)}
    tri_ids = sorted(tuple(sorted(sid[p] for p in tri)) for tri in keep)
    seg_arr = np.array([(i, j) if v4[i][-1] <= v4[j][-1] else (j, i) for (i, j) in segs], dtype=np.int64).reshape(-1, 2)
    return seg_arr, np.array(tri_ids, dtype=np.int64).reshape(-1, 3)


# ---------------------------------------------------------------------------------------------------
# Seeded tracking in 4D: GridContour4D inherits the tracker of tetrahedral.py:396-469 with the 16-corner
# box and the 80 offsets of pentatopes.py:32-39.  Pinned by tests/golden/seeded4d_*.npz (runs of the unmodified reference).
# ---------------------------------------------------------------------------------------------------
OFFSETS4D = np.array([(i, j, k, l) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) for l in (-1, 0, 1)
                      if i != 0 or j != 0 or k != 0 or l != 0], dtype=np.int64)


def initial_voxels(field, value, end_points):
    """tetrahedral.py:396-441 find_initial_voxels on a 4D array of samples (hypervoxels whose 16 samples are not all
    inside the array count as "not border").  Returns the set of start hypervoxel origins (tuples)."""
    field = np.asarray(field)
    value = float(value)
    act = active_cells(field, value)
    visited, new_voxels = set(), set()

    def border(p):
        return bool(all(0 <= int(p[a]) < act.shape[a] for a in range(4)) and act[tuple(int(x) for x in p)])

    for low_point, high_point in np.asarray(end_points, dtype=np.int64).reshape(-1, 2, 4):
        low_value, high_value = float(field[tuple(low_point)]), float(field[tuple(high_point)])
        if low_value > value or high_value < value:
            low_point, low_value, high_point, high_value = high_point, high_value, low_point, low_value
        assert low_value <= value and high_value >= value
        while np.any(np.abs(low_point - high_point) > 1):
            mid_point = (low_point + high_point) // 2
            if float(field[tuple(mid_point)]) < value:
                low_point = mid_point
            else:
                high_point = mid_point
        for point in (low_point, high_point):
            tpoint = tuple(int(x) for x in point)
            if tpoint in visited:
                continue
            visited.add(tpoint)
            if border(point):
                new_voxels.add(tpoint)
                continue
            for offset_point in OFFSETS4D + point.reshape(1, 4):
                toffset = tuple(int(x) for x in offset_point)
                if toffset in visited:
                    continue
                visited.add(toffset)
                if border(offset_point):
                    new_voxels.add(toffset)
                    break
    return new_voxels


def flood_fill(field, value, start_voxels):
    "tetrahedral.py:443-469 until nothing is new: border hypervoxels 80-connected, through border hypervoxels, to a start."
    from scipy import ndimage
    act = active_cells(np.asarray(field), float(value))
    lab, _ = ndimage.label(act, structure=np.ones((3, 3, 3, 3), dtype=bool))
    want = set(int(lab[tuple(v)]) for v in start_voxels if all(0 <= v[a] < act.shape[a] for a in range(4)) and act[tuple(v)])
    return np.isin(lab, sorted(want)) & act if want else np.zeros_like(act)


def extract_seeded(field, value, end_points, geom_dtype=np.float64):
    "extract() restricted to the hypervoxels the tracker reaches from the seed segments; keys renumbered."
    r = extract(field, value, geom_dtype)
    mask = flood_fill(field, value, initial_voxels(field, value, end_points))
    keep_t = mask.reshape(-1)[r["tet_cell"]] if len(r["tet_cell"]) else np.zeros(0, dtype=bool)
    tets = r["tets"][keep_t]
    used = np.unique(tets)
    remap = -np.ones(len(r["keys"]), dtype=np.int64)
    remap[used] = np.arange(len(used))
    return dict(keys=r["keys"][used], lowmin=r["lowmin"][used], pos=r["pos"][used], tets=remap[tets].astype(np.int32),
                tet_keys=r["tet_keys"][keep_t], voxels=mask)
