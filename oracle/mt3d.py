"""ORACLE (test infrastructure, not product code) -- 3D marching tetrahedra.

CPU restatement in numpy of the reference's 3D hot path, as a FULL SCAN over every
voxel (with `segment_endpoints=[]` + `search_for_endpoints()`
the reference's seeded tracking visits exactly the voxels a full scan finds).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (contourist_b200/) never does.

Parity status: PINNED.  Checked against golden vectors produced by running the
unmodified reference in the build container (tests/golden/make_golden.py ->
tests/golden/*.npz; tests/test_oracle_golden_3d.py, tests/test_oracle_seeded_3d.py: raw
extraction, the full C1 run, oriented final meshes, seeded tracker runs) and against the
reference's own known-answer test (contourist/test/test_tetrahedral.py:29-37).  Normals are NOT
computed by the reference (html_demo.py:142 "normals": []), so `normals()` below
is parity-UNPINNED: it is the definition, not a restatement.

Reference lines restated here (all under /root/reference/contourist/):
  tetrahedral.py:20-39    CUBE / TETRAHEDRA constants            -> CUBE, TETS
  tetrahedral.py:383-394  border_voxel (active-voxel predicate)  -> active_cells
  tetrahedral.py:561-595  enumerate_tetrahedron_triangles        -> tet_cases / extract
  tetrahedral.py:471-487,506-511 contour_pair_interpolation      -> interpolate
  (tetrahedral.py:190-215,353-375 quantize / tiny and surface_geometry.py:14-50 clean depend on CPython dict / set
   order and are not restated: SURVEY.md 8 "P-final, second tier")
  surface_geometry.py:52-140  orient_triangles                   -> orient
  tetrahedral.py:396-441  find_initial_voxels                    -> initial_voxels
  tetrahedral.py:443-469  expand_voxels / in_range               -> flood_fill, extract_seeded
  grid_field.py:64-84     find_contour_crossing_grid_segments    -> crossing_segments
  grid_field.py:89-93     from_grid_coordinates                  -> to_world

Conventions shared with the CUDA engine (include/contourist_b200.h):
  * field[i, j, k] = f at grid point (i, j, k); C order, k contiguous.
  * voxel origin o in [0, n-2]^3 ("cells"); linear point index lin = (i*n1 + j)*n2 + k.
  * edge direction code d = di*4 + dj*2 + dk in 1..7 (the 7 Kuhn edge directions);
    edge key = lin(min endpoint)*8 + d.  `lowmin` = 1 when the min endpoint is the
    LOW end (f < value), i.e. the reference's key is (min, max); else (max, min).
  * per-tet case code = 4-bit low mask (bit0 A, bit1 H, bit2 x_k, bit3 y_k for tet
    [A, H, x_k, y_k]) | 16 if the tet is skipped by np.allclose(values, value);
    cell code = sum(code_k << 5k), k = 0..5.
  * 2-2 quads are split with the rule "what the reference would do if `set` were
    `sorted`": [a, b] = sorted(least), [c, d] = sorted(most); triangles
    (ad, ac, bc) and (ad, bd, bc)  (tetrahedral.py:592-595).
"""
import numpy as np

RTOL = 1e-5
ATOL = 1e-8

# tetrahedral.py:20-29
CUBE = np.array([(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1),
                 (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)], dtype=np.int64)
A, B, C, D, E, F, G, H = range(8)       # corner number = di*4 + dj*2 + dk
# tetrahedral.py:32-39
TETS = np.array([[A, H, B, D], [A, H, D, C], [A, H, C, G],
                 [A, H, G, E], [A, H, E, F], [A, H, F, B]], dtype=np.int64)


def _corner_views(field):
    """8 views of shape (n0-1, n1-1, n2-1): value at each cube corner."""
    n0, n1, n2 = field.shape
    out = []
    for (di, dj, dk) in CUBE:
        out.append(field[di:n0 - 1 + di, dj:n1 - 1 + dj, dk:n2 - 1 + dk])
    return out


def near_f(fvals, value):
    """np.allclose(value, fvals) element test: |value - f| <= atol + rtol*|f|  (tetrahedral.py:391)."""
    fv = fvals.astype(np.float64)
    return np.abs(value - fv) <= ATOL + RTOL * np.abs(fv)


def near_v(fvals, value):
    """np.allclose(fvals, value) element test: |f - value| <= atol + rtol*|value|  (tetrahedral.py:576)."""
    fv = fvals.astype(np.float64)
    return np.abs(fv - value) <= ATOL + RTOL * abs(float(value))


def active_cells(field, value):
    """Boolean (n0-1, n1-1, n2-1): tetrahedral.py:383-394 border_voxel on every in-range voxel."""
    value = float(value)
    cs = _corner_views(field)
    mn = cs[0].astype(np.float64)
    mx = cs[0].astype(np.float64)
    allnear = near_f(cs[0], value)
    for c in cs[1:]:
        c64 = c.astype(np.float64)
        mn = np.minimum(mn, c64)
        mx = np.maximum(mx, c64)
        allnear &= near_f(c, value)
    return (~allnear) & (mn <= value) & (mx >= value)


def tet_cases(field, value):
    """uint32 (n0-1, n1-1, n2-1) packed per-cell case codes (see module docstring).

    tetrahedral.py:561-578: low = f < value (strict), high otherwise; a tet is skipped when
    either side is empty or np.allclose(values, value)."""
    value = float(value)
    cs = _corner_views(field)
    low = [c.astype(np.float64) < value for c in cs]
    nv = [near_v(c, value) for c in cs]
    code = np.zeros(cs[0].shape, dtype=np.uint32)
    for k, tet in enumerate(TETS):
        m = np.zeros(cs[0].shape, dtype=np.uint32)
        alln = np.ones(cs[0].shape, dtype=bool)
        for b, corner in enumerate(tet):
            m |= low[corner].astype(np.uint32) << b
            alln &= nv[corner]
        m |= alln.astype(np.uint32) << 4
        code |= m << np.uint32(5 * k)
    return code


def tet_emits(code5):
    """True where a 5-bit tet code produces triangles."""
    m = code5 & 15
    return (m != 0) & (m != 15) & ((code5 & 16) == 0)


def extract(field, value, geom_dtype=np.float64):
    """Full-scan raw extraction (before quantize/tiny/clean/orient).

    Returns dict with
      cells      int64 [Nc]      linear index (over the (n0-1,n1-1,n2-1) cell grid) of emitting cells, sorted
      codes      uint32 [Nc]     packed case codes of those cells
      keys       uint64 [V]      sorted unique edge keys
      lowmin     uint8 [V]       orientation of the reference key (1: (min,max), 0: (max,min))
      pos        geom_dtype [V,3] grid-coordinate position per key (tetrahedral.py:482-487)
      tri_keys   uint64 [T,3]    triangles as key triples, in (cell, tet, split) order
      tris       int32 [T,3]     same as indices into keys
      tri_cell   int64 [T]       owning cell (linear cell index)
      tri_tet    uint8 [T]       tet number 0..5
    """
    field = np.asarray(field)
    value = float(value)
    n0, n1, n2 = field.shape
    act = active_cells(field, value)
    code = tet_cases(field, value)
    emits_any = np.zeros(act.shape, dtype=bool)
    for k in range(6):
        emits_any |= tet_emits((code >> np.uint32(5 * k)) & np.uint32(31))
    sel = act & emits_any
    ci, cj, ck = np.nonzero(sel)
    cell_lin = (ci * (n1 - 1) + cj) * (n2 - 1) + ck
    codes = code[sel]
    tri_keys = []
    tri_cell = []
    tri_tet = []
    tri_ord = []
    lin0 = (ci * n1 + cj) * n2 + ck                      # lin of corner A
    corner_lin = [lin0 + (int(d[0]) * n1 + int(d[1])) * n2 + int(d[2]) for d in CUBE]

    def key_of(ca, cb):
        # ca, cb corner numbers (python ints); min endpoint = smaller corner number
        # (corner number order == lexicographic point order inside one cube).
        lo_c, hi_c = (ca, cb) if ca < cb else (cb, ca)
        d = hi_c - lo_c           # corner numbers are bit patterns; hi has all bits of lo here
        return (corner_lin[lo_c].astype(np.uint64) << np.uint64(3)) | np.uint64(d)

    for k, tet in enumerate(TETS):
        c5 = (codes >> np.uint32(5 * k)) & np.uint32(31)
        em = tet_emits(c5)
        for mask in range(1, 15):
            rows = np.nonzero(em & ((c5 & 15) == mask))[0]
            if rows.size == 0:
                continue
            lows = sorted(int(tet[b]) for b in range(4) if (mask >> b) & 1)
            highs = sorted(int(tet[b]) for b in range(4) if not (mask >> b) & 1)
            least, most = lows, highs
            if len(least) > len(most):
                least, most = most, least
            if len(least) == 1:
                a = least[0]
                b_, c_, d_ = most
                tl = [((a, b_), (a, c_), (a, d_))]
            else:
                a, b_ = least
                c_, d_ = most
                tl = [((a, d_), (a, c_), (b_, c_)), ((a, d_), (b_, d_), (b_, c_))]
            for s, tri in enumerate(tl):
                ks = np.stack([key_of(u, w)[rows] for (u, w) in tri], axis=1)
                tri_keys.append(ks)
                tri_cell.append(cell_lin[rows])
                tri_tet.append(np.full(rows.size, k, dtype=np.uint8))
                tri_ord.append(cell_lin[rows] * 12 + k * 2 + s)
    if tri_keys:
        tri_keys = np.concatenate(tri_keys)
        tri_cell = np.concatenate(tri_cell)
        tri_tet = np.concatenate(tri_tet)
        order = np.argsort(np.concatenate(tri_ord), kind="stable")
        tri_keys, tri_cell, tri_tet = tri_keys[order], tri_cell[order], tri_tet[order]
    else:
        tri_keys = np.zeros((0, 3), dtype=np.uint64)
        tri_cell = np.zeros((0,), dtype=np.int64)
        tri_tet = np.zeros((0,), dtype=np.uint8)
    keys = np.unique(tri_keys)
    tris = np.searchsorted(keys, tri_keys).astype(np.int32)
    lowmin, pos = interpolate(field, value, keys, geom_dtype)
    return dict(cells=cell_lin.astype(np.int64), codes=codes, keys=keys, lowmin=lowmin, pos=pos,
                tri_keys=tri_keys, tris=tris, tri_cell=tri_cell, tri_tet=tri_tet)


def key_points(keys, shape):
    """Decode keys -> (pmin[V,3], pmax[V,3]) integer grid points."""
    n0, n1, n2 = shape
    lin = (keys >> np.uint64(3)).astype(np.int64)
    d = (keys & np.uint64(7)).astype(np.int64)
    i = lin // (n1 * n2)
    j = (lin // n2) % n1
    k = lin % n2
    pmin = np.stack([i, j, k], axis=1)
    pmax = pmin + np.stack([(d >> 2) & 1, (d >> 1) & 1, d & 1], axis=1)
    return pmin, pmax


def interpolate(field, value, keys, geom_dtype=np.float64):
    """tetrahedral.py:471-487: key oriented (low, high) by VALUE (swap iff flow > fhigh);
    ratio = (z - flow)/(fhigh - flow), or 0.5 when np.allclose(fhigh - flow, 0);
    x = low + ratio*(high - low) in grid coordinates.
    geom_dtype float32 mirrors the engine's fp32 geometry mode (same formula in float)."""
    shape = field.shape
    pmin, pmax = key_points(keys, shape)
    fmin = field[pmin[:, 0], pmin[:, 1], pmin[:, 2]]
    fmax = field[pmax[:, 0], pmax[:, 1], pmax[:, 2]]
    gd = np.dtype(geom_dtype)
    fmin = fmin.astype(gd)
    fmax = fmax.astype(gd)
    swap = fmin > fmax                       # then low point is pmax
    lowmin = (~swap).astype(np.uint8)
    flow = np.where(swap, fmax, fmin)
    fhigh = np.where(swap, fmin, fmax)
    plow = np.where(swap[:, None], pmax, pmin).astype(gd)
    phigh = np.where(swap[:, None], pmin, pmax).astype(gd)
    z = gd.type(value)
    den = fhigh - flow
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (z - flow) / den
    ratio = np.where(np.abs(den.astype(np.float64)) <= ATOL, gd.type(0.5), ratio).astype(gd)
    pos = plow + ratio[:, None] * (phigh - plow)
    return lowmin, pos.astype(gd)


def gradient(field):
    """fp64 central differences (one-sided at the domain boundary), grid units. [n0,n1,n2,3]"""
    f = field.astype(np.float64)
    g = np.zeros(f.shape + (3,), dtype=np.float64)
    for ax in range(3):
        sl = [slice(None)] * 3
        lo = list(sl); hi = list(sl); mid = list(sl)
        mid[ax] = slice(1, -1); lo[ax] = slice(0, -2); hi[ax] = slice(2, None)
        g[tuple(mid) + (ax,)] = 0.5 * (f[tuple(hi)] - f[tuple(lo)])
        first = list(sl); first[ax] = 0; second = list(sl); second[ax] = 1
        g[tuple(first) + (ax,)] = f[tuple(second)] - f[tuple(first)]
        last = list(sl); last[ax] = -1; prev = list(sl); prev[ax] = -2
        g[tuple(last) + (ax,)] = f[tuple(last)] - f[tuple(prev)]
    return g


def normals(field, value, keys, geom_dtype=np.float64, delta=(1.0, 1.0, 1.0)):
    """Gradient normals (parity UNPINNED by the reference -- engine definition):
    n = normalize( (g(low) + ratio*(g(high) - g(low))) / delta ), g = central-difference gradient
    (one-sided on the boundary), same `ratio` as the position. Zero gradient -> (0,0,0)."""
    shape = field.shape
    gd = np.dtype(geom_dtype)
    pmin, pmax = key_points(keys, shape)
    f = field.astype(gd)

    def grad_at(p):
        out = np.zeros((p.shape[0], 3), dtype=gd)
        for ax in range(3):
            n = shape[ax]
            pl = p.copy(); ph = p.copy()
            pl[:, ax] = np.maximum(p[:, ax] - 1, 0)
            ph[:, ax] = np.minimum(p[:, ax] + 1, n - 1)
            scale = np.where((p[:, ax] == 0) | (p[:, ax] == n - 1), gd.type(1.0), gd.type(0.5))
            out[:, ax] = (f[ph[:, 0], ph[:, 1], ph[:, 2]] - f[pl[:, 0], pl[:, 1], pl[:, 2]]) * scale
        return out

    fmin = f[pmin[:, 0], pmin[:, 1], pmin[:, 2]]
    fmax = f[pmax[:, 0], pmax[:, 1], pmax[:, 2]]
    swap = fmin > fmax
    flow = np.where(swap, fmax, fmin)
    fhigh = np.where(swap, fmin, fmax)
    den = fhigh - flow
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (gd.type(value) - flow) / den
    ratio = np.where(np.abs(den.astype(np.float64)) <= ATOL, gd.type(0.5), ratio).astype(gd)
    gmin = grad_at(pmin)
    gmax = grad_at(pmax)
    glow = np.where(swap[:, None], gmax, gmin)
    ghigh = np.where(swap[:, None], gmin, gmax)
    g = glow + ratio[:, None] * (ghigh - glow)
    g = g / np.asarray(delta, dtype=gd)[None, :]
    nrm = np.sqrt((g * g).sum(axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(nrm[:, None] > 0, g / nrm[:, None], gd.type(0))
    return out.astype(gd)


def crossing_segments(field, value, count_only=False):
    """grid_field.py:64-84 with skip=1 over an array holding samples 0..N (N+1 per axis):
    for every v0 in [0,N-1]^3 and v1 = v0 + {0,1}^3, v1 != v0: seed iff (f0-v)*(f1-v) < 0.
    Returns (maxf, minf, count) or the (v0_lin*8 + d) list."""
    f = field.astype(np.float64) - float(value)
    n0, n1, n2 = field.shape
    base = f[:n0 - 1, :n1 - 1, :n2 - 1]
    total = 0
    out = []
    for c in range(1, 8):
        di, dj, dk = CUBE[c]
        nb = f[di:n0 - 1 + di, dj:n1 - 1 + dj, dk:n2 - 1 + dk]
        hit = (base * nb) < 0
        if count_only:
            total += int(hit.sum())
        else:
            i, j, k = np.nonzero(hit)
            out.append((((i * n1 + j) * n2 + k).astype(np.uint64) << np.uint64(3)) | np.uint64(c))
    fa = field.astype(np.float64)
    if count_only:
        return fa.max(), fa.min(), total
    return fa.max(), fa.min(), np.sort(np.concatenate(out)) if out else np.zeros(0, np.uint64)


def to_world(pos, mins, delta):
    """grid_field.py:89-93."""
    return np.asarray(pos, dtype=np.float64) * np.asarray(delta, dtype=np.float64) + np.asarray(mins, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# Stage 4b: the reference's serial post-processing, restated with the engine's deterministic
# tie-breaks (SURVEY.md section 7 hard part 1: the reference's own choices depend on CPython set/dict order).
# ----------------------------------------------------------------------------------------------

def orient(verts, tris):
    """surface_geometry.py:52-140 with compatible_triangle_test = always True.
    Per connected component (edge adjacency): seed = the triangle at the max-x vertex with the largest
    |cross.x| (ties: reference keeps the LAST in set order; here the largest sorted triple), wound so
    cross(a-b, a-c).x > 0 (>= 0 kept as is), propagated across shared edges (DFS, LIFO).
    Returns sorted list of oriented (i, j, k)."""
    verts = np.asarray(verts, dtype=np.float64)
    tl = [tuple(int(x) for x in row) for row in tris]
    unoriented = set(frozenset(t) for t in tl if len(set(t)) == 3)
    seg2tri = {}
    pt2tri = {}
    for t in sorted(unoriented, key=lambda s: tuple(sorted(s))):
        a, b, c = sorted(t)
        for i in (a, b, c):
            pt2tri.setdefault(i, []).append(t)
        for e in ((a, b), (b, c), (a, c)):
            seg2tri.setdefault(frozenset(e), []).append(t)
    orientations = {}
    while unoriented:
        vs = set()
        for t in unoriented:
            vs.update(t)
        max_x, max_index = max((verts[i][0], i) for i in vs)
        cands = [t for t in pt2tri[max_index] if t in unoriented]
        init = None
        maxdot = 0.0
        for t in cands:
            a, b, c = (verts[i] for i in sorted(t))
            dotx = np.cross(a - b, a - c)[0]
            if abs(dotx) >= abs(maxdot):
                maxdot = dotx
                init = t
        o = tuple(sorted(init))
        a, b, c = (verts[i] for i in o)
        if np.cross(a - b, a - c)[0] < 0:
            o = tuple(reversed(o))
        stack = [(init, o)]
        while stack:
            t, o = stack.pop()
            orientations[t] = o
            unoriented.discard(t)
            a, b, c = o
            for (i1, i2) in ((c, b), (b, a), (a, c)):
                e = frozenset((i1, i2))
                for t2 in seg2tri[e]:
                    if t2 != t and t2 not in orientations:
                        (i3,) = t2 - e
                        stack.append((t2, (i1, i2, i3)))
    return sorted(orientations.values())


# ---------------------------------------------------------------------------------------------------
# Seeded tracking: what the reference returns for explicit seed segments
# ---------------------------------------------------------------------------------------------------
OFFSETS = np.array([(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)
                    if i != 0 or j != 0 or k != 0], dtype=np.int64)       # tetrahedral.py:41-47


def initial_voxels(field, value, end_points):
    """tetrahedral.py:396-441 find_initial_voxels on an array of samples.  A voxel whose 8 samples are not all inside
    the array cannot be evaluated and counts as "not border" (the reference, through a callable, may call it border:
    its out-of-range leak voxels).  Returns the set of start voxel origins (tuples)."""
    field = np.asarray(field)
    value = float(value)
    act = active_cells(field, value)
    visited = set()
    new_voxels = set()

    def border(p):
        return bool(all(0 <= int(p[a]) < act.shape[a] for a in range(3)) and act[tuple(int(x) for x in p)])

    for low_point, high_point in np.asarray(end_points, dtype=np.int64).reshape(-1, 2, 3):
        low_value = float(field[tuple(low_point)])
        high_value = float(field[tuple(high_point)])
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
            for offset_point in OFFSETS + point.reshape(1, 3):
                toffset = tuple(int(x) for x in offset_point)
                if toffset in visited:
                    continue
                visited.add(toffset)
                if border(offset_point):
                    new_voxels.add(toffset)
                    break
    return new_voxels


def flood_fill(field, value, start_voxels):
    """tetrahedral.py:443-469 expand_voxels / in_range until nothing is new: the border voxels (active_cells) that are
    26-connected, through border voxels, to a start voxel.  Boolean (n0-1, n1-1, n2-1)."""
    from scipy import ndimage
    act = active_cells(np.asarray(field), float(value))
    lab, _ = ndimage.label(act, structure=np.ones((3, 3, 3), dtype=bool))
    want = set()
    for v in start_voxels:
        if all(0 <= v[a] < act.shape[a] for a in range(3)) and act[tuple(v)]:
            want.add(int(lab[tuple(v)]))
    return np.isin(lab, sorted(want)) & act if want else np.zeros_like(act)


def extract_seeded(field, value, end_points, geom_dtype=np.float64):
    """extract() restricted to the voxels the tracker reaches from the seed segments; keys renumbered.
    Returns the same dict as extract() plus 'voxels' (bool mask of the reached border voxels)."""
    r = extract(field, value, geom_dtype)
    mask = flood_fill(field, value, initial_voxels(field, value, end_points))
    keep_t = mask.reshape(-1)[r["tri_cell"]] if len(r["tri_cell"]) else np.zeros(0, dtype=bool)
    tris = r["tris"][keep_t]
    used = np.unique(tris)
    remap = -np.ones(len(r["keys"]), dtype=np.int64)
    remap[used] = np.arange(len(used))
    keep_c = mask.reshape(-1)[r["cells"]]
    return dict(cells=r["cells"][keep_c], codes=r["codes"][keep_c], keys=r["keys"][used], lowmin=r["lowmin"][used],
                pos=r["pos"][used], tri_keys=r["tri_keys"][keep_t], tris=remap[tris].astype(np.int32),
                tri_cell=r["tri_cell"][keep_t], tri_tet=r["tri_tet"][keep_t], voxels=mask)
