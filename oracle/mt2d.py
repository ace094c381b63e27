"""ORACLE (test infrastructure, not product code) -- 2D marching triangles, multi-level.

numpy restatement of the reference's 2D contour path as a FULL SCAN.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Parity status: PINNED against golden vectors made by running the unmodified reference
(tests/golden/make_golden.py -> mt2d_*.npz, seeded2d_*.npz) and the reference's own known answers
(contourist/test/test_triangulated.py:20-42,79-106).

Reference lines restated (under /root/reference/contourist/):
  triangulated.py:10-14    adjacent_offsets: the (1,1)-diagonal triangulated grid
  triangulated.py:66-77    adjacent_pairs
  triangulated.py:340-362  in_range / contour_pair_interpolation: key (low, high) exists iff
                           f(low) <= z <= f(high) (inclusive both sides); ratio 0.5 when the
                           denominator is allclose to 0, else (z - flow)/(fhigh - flow)
  triangulated.py:295-305  find_adjacencies: two keys are linked iff they are adjacent_pairs of each other
  triangulated.py:307-338  find_initial_contour_pairs / expand_contour_pairs (seeded tracking) -> seeded_keys()
  triangulated.py:221-293  get_contour_sequences (polyline chaining)       -> polylines()
  multiple_2d_contour.py:50-75   level classification of grid edges         -> (same predicate per level)
  multiple_2d_contour.py:100-108 Linear2DContour.get_values                 -> linear_levels()
  multiple_2d_contour.py:91-98   Percentile2DContour.get_values             -> percentile_levels()

Conventions shared with the CUDA engine:
  field[i, j] = f at grid point (i, j); points in range are 0 <= i < n0, 0 <= j < n1.
  edge direction d = di*2 + dj in {1: (0,1), 2: (1,0), 3: (1,1)}; owner = component-wise min endpoint;
  key = ((lin(owner)*4 + d) << 1) | lowmin, lowmin = 1 when the low end is the owner (reference key (p, q)),
  0 for the reference key (q, p).  Both exist when f(p) == f(q) == z.
  grid triangles of square (i, j): T0 = {(i,j), (i+1,j), (i+1,j+1)}, T1 = {(i,j), (i,j+1), (i+1,j+1)}.
  A segment joins the two keys of one triangle that share their low (or their high) endpoint.
"""
import numpy as np

ATOL = 1e-8


def key_of(lin_owner, d, lowmin):
    return ((lin_owner.astype(np.uint64) * np.uint64(4) + np.uint64(d)) << np.uint64(1)) | lowmin.astype(np.uint64)


def _interp(flow, fhigh, plow, phigh, z, gd):
    den = fhigh - flow
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (gd.type(z) - flow) / den
    ratio = np.where(np.abs(den.astype(np.float64)) <= ATOL, gd.type(0.5), ratio).astype(gd)
    return (plow + ratio[:, None] * (phigh - plow)).astype(gd)


def extract_level(field, z, geom_dtype=np.float64):
    """All keys and segments of one level.

    Returns dict: keys uint64 [V] sorted unique, pos [V,2] grid coordinates,
                  seg_keys uint64 [S,2] (each row sorted), unique rows, lexicographically sorted."""
    f = np.asarray(field)
    gd = np.dtype(geom_dtype)
    n0, n1 = f.shape
    z = float(z)
    f64 = f.astype(np.float64)
    lin = (np.arange(n0)[:, None] * n1 + np.arange(n1)[None, :])
    keys, pos = [], []
    for d, (di, dj) in ((1, (0, 1)), (2, (1, 0)), (3, (1, 1))):
        fp = f64[:n0 - di, :n1 - dj]
        fq = f64[di:, dj:]
        lp = lin[:n0 - di, :n1 - dj]
        ii, jj = np.meshgrid(np.arange(n0 - di), np.arange(n1 - dj), indexing="ij")
        for lowmin in (1, 0):
            ex = (fp <= z) & (z <= fq) if lowmin else (fq <= z) & (z <= fp)
            sel = np.nonzero(ex)
            if sel[0].size == 0:
                continue
            P = np.stack([ii[sel], jj[sel]], axis=1)
            Q = P + np.array([di, dj])
            a, b = (fp[sel], fq[sel]) if lowmin else (fq[sel], fp[sel])
            pl, ph = (P, Q) if lowmin else (Q, P)
            keys.append(key_of(lp[sel], d, np.full(sel[0].size, lowmin)))
            pos.append(_interp(a.astype(gd), b.astype(gd), pl.astype(gd), ph.astype(gd), z, gd))
    if keys:
        keys = np.concatenate(keys)
        pos = np.concatenate(pos)
        o = np.argsort(keys)
        keys, pos = keys[o], pos[o]
    else:
        keys = np.zeros(0, np.uint64)
        pos = np.zeros((0, 2), gd)
    # segments: per triangle, per isolated vertex a (low-iso / high-iso)
    segs = []

    def edge_key(pa, pb, low_is_a):
        """key of the edge between integer point arrays pa, pb with low end pa (low_is_a) or pb."""
        pm = np.minimum(pa, pb)
        d = (np.maximum(pa, pb) - pm)
        dcode = d[:, 0] * 2 + d[:, 1]
        low = pa if low_is_a else pb
        lowmin = (low == pm).all(axis=1)
        return ((((pm[:, 0] * n1 + pm[:, 1]).astype(np.uint64) * np.uint64(4) + dcode.astype(np.uint64)) << np.uint64(1))
                | lowmin.astype(np.uint64))

    ii, jj = np.meshgrid(np.arange(n0 - 1), np.arange(n1 - 1), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    base = np.stack([ii, jj], axis=1)
    for tri in (((0, 0), (1, 0), (1, 1)), ((0, 0), (0, 1), (1, 1))):
        pts = [base + np.array(o) for o in tri]
        fv = [f64[p[:, 0], p[:, 1]] for p in pts]
        for a in range(3):
            b, c = [x for x in range(3) if x != a]
            lo = (fv[a] <= z) & (z <= fv[b]) & (z <= fv[c])
            hi = (fv[b] <= z) & (fv[c] <= z) & (z <= fv[a])
            for cond, low_is_a in ((lo, True), (hi, False)):
                s = np.nonzero(cond)[0]
                if s.size == 0:
                    continue
                k1 = edge_key(pts[a][s], pts[b][s], low_is_a)
                k2 = edge_key(pts[a][s], pts[c][s], low_is_a)
                segs.append(np.stack([np.minimum(k1, k2), np.maximum(k1, k2)], axis=1))
    if segs:
        segs = np.unique(np.concatenate(segs), axis=0)
    else:
        segs = np.zeros((0, 2), np.uint64)
    return dict(keys=keys, pos=pos, seg_keys=segs)


def decode_key(keys, shape):
    n0, n1 = shape
    lowmin = (keys & np.uint64(1)).astype(np.int64)
    kd = keys >> np.uint64(1)
    d = (kd & np.uint64(3)).astype(np.int64)
    lin = (kd >> np.uint64(2)).astype(np.int64)
    p = np.stack([lin // n1, lin % n1], axis=1)
    q = p + np.stack([(d >> 1) & 1, d & 1], axis=1)
    low = np.where(lowmin[:, None] == 1, p, q)
    high = np.where(lowmin[:, None] == 1, q, p)
    return low, high


def linear_levels(field, breakpoints):
    """multiple_2d_contour.py:100-108: offset = (max-min)/breakpoints; values = offset*i, i=1..breakpoints-1
    (note: NOT shifted by the minimum -- restated as is)."""
    s = np.asarray(field, dtype=np.float64)
    mn, mx = s.min(), s.max()
    offset = (mx - mn) * (1.0 / breakpoints)
    return [offset * i for i in range(1, breakpoints)]


def percentile_levels(field, breakpoints):
    """multiple_2d_contour.py:91-98."""
    s = np.sort(np.asarray(field, dtype=np.float64).flatten())
    n = s.shape[0]
    skip = int(n / breakpoints)
    return [s[i] for i in range(skip, n, skip)]


def polylines(keys, pos, seg_keys):
    """Chain segments into polylines: restates triangulated.py:236-293 for the generic case (every key has
    at most two neighbours).  Deterministic: open polylines start at their smaller end key, closed ones at
    their smallest key and run towards the smaller neighbour.  Consecutive (allclose) duplicate points are
    dropped (triangulated.py:269).  Returns [(closed, keys_in_order uint64[], points[k,2])]."""
    idx = {int(k): n for n, k in enumerate(keys)}
    nbr = {int(k): [] for k in keys}
    for a, b in seg_keys:
        a, b = int(a), int(b)
        if a != b:
            nbr[a].append(b)
            nbr[b].append(a)
    visited = set()
    out = []
    ends = sorted(k for k in nbr if len(nbr[k]) < 2)
    order = ends + sorted(k for k in nbr if len(nbr[k]) >= 2)
    for start in order:
        if start in visited:
            continue
        closed = len(nbr[start]) >= 2
        chain = [start]
        visited.add(start)
        cur = start
        while True:
            nxt = [k for k in sorted(nbr[cur]) if k not in visited]
            if not nxt:
                break
            cur = nxt[0]
            visited.add(cur)
            chain.append(cur)
        pts = []
        for k in chain:
            p = pos[idx[k]]
            if not pts or not np.allclose(pts[-1], p):
                pts.append(p)
        pts = np.array(pts)
        if len(pts) > 1 and np.allclose(pts[0], pts[-1]):
            closed = True
        out.append((closed, np.array(chain, dtype=np.uint64), pts))
    return out


# ---------------------------------------------------------------------------------------------------
# Seeded tracking: what the reference returns for explicit seed segments
# ---------------------------------------------------------------------------------------------------
def seeded_keys(field, z, end_points):
    """triangulated.py:307-338 find_initial_contour_pairs + expand_contour_pairs until nothing is new, on an array:
    bisect every seed segment down to adjacent points (a pair "exists" iff f(low) <= z <= f(high)), start from all
    contour pairs that have the low point as low end or the high point as high end, and keep adding the pairs that
    share their low end (as low) or their high end (as high) with a pair already found.  Returns the sorted keys."""
    f = np.asarray(field).astype(np.float64)
    z = float(z)
    r = extract_level(field, z)
    keys = r["keys"]
    low, high = decode_key(keys, f.shape)
    n1 = f.shape[1]
    by_low, by_high = {}, {}
    for q, (lo, hi) in enumerate(zip(low[:, 0] * n1 + low[:, 1], high[:, 0] * n1 + high[:, 1])):
        by_low.setdefault(int(lo), []).append(q)
        by_high.setdefault(int(hi), []).append(q)

    def exists(lo, hi):
        return f[tuple(lo)] <= z <= f[tuple(hi)]

    found, horizon = set(), set()
    for low_point, high_point in np.asarray(end_points, dtype=np.int64).reshape(-1, 2, 2):
        if not exists(low_point, high_point):
            low_point, high_point = high_point, low_point
            assert exists(low_point, high_point), "bad end points"
        while np.any(np.abs(low_point - high_point) > 1):
            mid_point = (low_point + high_point) // 2
            if exists(low_point, mid_point):
                high_point = mid_point
            else:
                assert exists(mid_point, high_point)
                low_point = mid_point
        new = set(by_low.get(int(low_point[0] * n1 + low_point[1]), [])) | \
            set(by_high.get(int(high_point[0] * n1 + high_point[1]), []))
        assert new
        found |= new
        horizon |= new
    while horizon:
        nxt = set()
        for q in horizon:
            nxt.update(by_low[int(low[q, 0] * n1 + low[q, 1])])
            nxt.update(by_high[int(high[q, 0] * n1 + high[q, 1])])
        horizon = nxt - found
        found |= nxt
    return keys[sorted(found)]


def extract_level_seeded(field, z, end_points, geom_dtype=np.float64):
    """extract_level restricted to the pairs the tracker reaches from the seed segments."""
    r = extract_level(field, z, geom_dtype)
    keep = np.isin(r["keys"], seeded_keys(field, z, end_points))
    segs = r["seg_keys"]
    sk = np.isin(segs[:, 0], r["keys"][keep]) if len(segs) else np.zeros(0, dtype=bool)
    return dict(keys=r["keys"][keep], pos=r["pos"][keep], seg_keys=segs[sk])
