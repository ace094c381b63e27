"""Generate contourist_b200/csrc/tables.h: the combinatorial lookup tables of the engine.

Derived from the reference's constants and case rules (paths under /root/reference/contourist/):
  tetrahedral.py:20-39     CUBE, TETRAHEDRA (6 Kuhn tets on diagonal A-H)
  tetrahedral.py:579-595   1-vs-3 -> triangle {ab,ac,ad}; 2-vs-2 -> {ad,ac,bc},{ad,bd,bc}
  pentatopes.py:15-30      PENTATOPES (24 Kuhn pentatopes = axis permutations), HYPERCUBE
  pentatopes.py:241-291    1-vs-4 -> tet {ab,ac,ad,ae}; 2-vs-3 -> {ac,be,ad,bd},{ac,be,ad,ae},{ac,be,bd,bc}
  triangulated.py:10-14    the (1,1)-diagonal triangulated 2D grid
The reference unpacks Python sets ([a, b] = leastpoints); the engine's deterministic rule is
"as if the sets were sorted lexicographically" (hard part 1).
3D triangles are additionally wound so that the geometric normal points to the HIGH side (+gradient).

Run:  python tools/gen_tables.py   (rewrites the header; output is committed)
"""
import itertools
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "contourist_b200", "csrc", "tables.h")

# ---------------------------------------------------------------- 3D
A, B, C, D, E, F, G, H = range(8)         # corner number = di*4 + dj*2 + dk
TETS = [[A, H, B, D], [A, H, D, C], [A, H, C, G], [A, H, G, E], [A, H, E, F], [A, H, F, B]]


def corner_xyz(c, nd):
    return [(c >> (nd - 1 - a)) & 1 for a in range(nd)]


def edges_of(simplices):
    es = set()
    for s in simplices:
        for u, w in itertools.combinations(s, 2):
            lo, hi = min(u, w), max(u, w)
            assert lo & hi == lo, "Kuhn edges are monotone"
            es.add((lo, hi - lo))
    return sorted(es)


EDGES3 = edges_of(TETS)
assert len(EDGES3) == 19
SLOT3 = {e: i for i, e in enumerate(EDGES3)}


def slot3(u, w):
    lo, hi = min(u, w), max(u, w)
    return SLOT3[(lo, hi - lo)]


def tri_table3():
    """[6][16] -> (ntri, 6 slots)"""
    tab = []
    for tet in TETS:
        row = []
        for mask in range(16):
            lows = sorted(tet[b] for b in range(4) if (mask >> b) & 1)
            highs = sorted(tet[b] for b in range(4) if not (mask >> b) & 1)
            if not lows or not highs:
                row.append((0, [255] * 6))
                continue
            least, most = lows, highs
            if len(least) > len(most):
                least, most = most, least
            if len(least) == 1:
                a = least[0]
                b_, c_, d_ = most
                tris = [[(a, b_), (a, c_), (a, d_)]]
            else:
                a, b_ = least
                c_, d_ = most
                tris = [[(a, d_), (a, c_), (b_, c_)], [(a, d_), (b_, d_), (b_, c_)]]
            # orientation: normal toward the high side
            P = np.array([corner_xyz(c, 3) for c in tet], dtype=float)
            fv = np.array([-1.0 if (mask >> b) & 1 else 1.0 for b in range(4)])
            M = np.hstack([P, np.ones((4, 1))])
            grad = np.linalg.solve(M, fv)[:3]
            slots = []
            for tri in tris:
                pts = [0.5 * (np.array(corner_xyz(u, 3), float) + np.array(corner_xyz(w, 3), float)) for u, w in tri]
                n = np.cross(pts[1] - pts[0], pts[2] - pts[0])
                s = float(n @ grad)
                assert abs(s) > 1e-9
                t = list(tri)
                if s < 0:
                    t[1], t[2] = t[2], t[1]
                slots += [slot3(u, w) for u, w in t]
            slots += [255] * (6 - len(slots))
            row.append((len(tris), slots))
        tab.append(row)
    return tab


def tetmask3():
    """[8 dirs][8 s] -> 6-bit mask of tets (in the cell at p - s) containing edge (s, s|d)."""
    out = [[0] * 8 for _ in range(8)]
    for d in range(1, 8):
        for s in range(8):
            if s & d:
                continue
            m = 0
            for k, tet in enumerate(TETS):
                if s in tet and (s | d) in tet:
                    m |= 1 << k
            out[d][s] = m
    return out


# ---------------------------------------------------------------- 4D
def pentatopes():
    out = []
    for perm in itertools.permutations(range(4)):
        v = [0, 0, 0, 0]
        verts = [0]
        for idx in perm:
            v[idx] = 1
            verts.append(v[0] * 8 + v[1] * 4 + v[2] * 2 + v[3])
        out.append(verts)
    return out


PENTS = pentatopes()
EDGES4 = edges_of(PENTS)
SLOT4 = {e: i for i, e in enumerate(EDGES4)}


def slot4(u, w):
    lo, hi = min(u, w), max(u, w)
    return SLOT4[(lo, hi - lo)]


def tet_table4():
    """[24][32] -> (ntet, 12 slots)"""
    tab = []
    for pent in PENTS:
        row = []
        for mask in range(32):
            lows = sorted(pent[b] for b in range(5) if (mask >> b) & 1)
            highs = sorted(pent[b] for b in range(5) if not (mask >> b) & 1)
            if not lows or not highs:
                row.append((0, [255] * 12))
                continue
            least, most = lows, highs
            if len(least) > len(most):
                least, most = most, least
            if len(least) == 1:
                a = least[0]
                b_, c_, d_, e_ = most
                tets = [[(a, b_), (a, c_), (a, d_), (a, e_)]]
            else:
                a, b_ = least
                c_, d_, e_ = most
                ac, ad, ae, bc, bd, be = (a, c_), (a, d_), (a, e_), (b_, c_), (b_, d_), (b_, e_)
                tets = [[ac, be, ad, bd], [ac, be, ad, ae], [ac, be, bd, bc]]
            slots = []
            for t in tets:
                slots += [slot4(u, w) for u, w in t]
            slots += [255] * (12 - len(slots))
            row.append((len(tets), slots))
        tab.append(row)
    return tab


def pentmask4():
    """[16 dirs][16 s] -> 24-bit mask of pentatopes (in the hypercube at p - s) containing edge (s, s|d)."""
    out = [[0] * 16 for _ in range(16)]
    for d in range(1, 16):
        for s in range(16):
            if s & d:
                continue
            m = 0
            for k, pent in enumerate(PENTS):
                if s in pent and (s | d) in pent:
                    m |= 1 << k
            out[d][s] = m
    return out


def carr(name, ctype, data, dims):
    flat = np.array(data).reshape(-1)
    body = ", ".join(str(int(x)) for x in flat)
    dim = "".join("[%d]" % d for d in dims)
    return "static const %s %s%s = { %s };\n" % (ctype, name, dim, body)


def main():
    t3 = tri_table3()
    t4 = tet_table4()
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_tables.py -- do not edit.\n")
        f.write("// Case tables of the marching-tetrahedra / pentatope engine; see the generator for the\n")
        f.write("// reference lines (tetrahedral.py:20-39,579-595; pentatopes.py:15-30,241-291) they restate.\n")
        f.write("#pragma once\n#include <stdint.h>\n\n")
        f.write("#define CTR_NEDGE3 19\n#define CTR_NEDGE4 %d\n\n" % len(EDGES4))
        f.write("// 3D: edge slot -> owner corner s (di*4+dj*2+dk) and direction d\n")
        f.write(carr("CTR_EDGE3_S_H", "uint8_t", [e[0] for e in EDGES3], [19]))
        f.write(carr("CTR_EDGE3_D_H", "uint8_t", [e[1] for e in EDGES3], [19]))
        f.write("// 3D: tet k = [A, H, x_k, y_k] corner numbers\n")
        f.write(carr("CTR_TET3_H", "uint8_t", TETS, [6, 4]))
        f.write("// 3D: [tet][4-bit low mask] -> triangle count, and 6 edge slots (2 triangles x 3), wound so the normal points to the high side\n")
        f.write(carr("CTR_TRI3_N_H", "uint8_t", [[c[0] for c in row] for row in t3], [6, 16]))
        f.write(carr("CTR_TRI3_E_H", "uint8_t", [[c[1] for c in row] for row in t3], [6, 16, 6]))
        f.write("// 3D: [d][s] -> tets of the cell at (p - s) that contain the edge p -> p+d\n")
        f.write(carr("CTR_TETMASK3_H", "uint8_t", tetmask3(), [8, 8]))
        f.write("\n// 4D: edge slot -> owner corner s (di*8+dj*4+dk*2+dl) and direction d\n")
        f.write(carr("CTR_EDGE4_S_H", "uint8_t", [e[0] for e in EDGES4], [len(EDGES4)]))
        f.write(carr("CTR_EDGE4_D_H", "uint8_t", [e[1] for e in EDGES4], [len(EDGES4)]))
        f.write("// 4D: pentatope k corner numbers (pentatopes.py:15-26 order)\n")
        f.write(carr("CTR_PENT4_H", "uint8_t", PENTS, [24, 5]))
        f.write("// 4D: [pentatope][5-bit low mask] -> tetrahedron count and 12 edge slots (3 tets x 4)\n")
        f.write(carr("CTR_TET4_N_H", "uint8_t", [[c[0] for c in row] for row in t4], [24, 32]))
        f.write(carr("CTR_TET4_E_H", "uint8_t", [[c[1] for c in row] for row in t4], [24, 32, 12]))
        f.write("// 4D: [d][s] -> pentatopes of the hypercube at (p - s) that contain the edge p -> p+d\n")
        f.write(carr("CTR_PENTMASK4_H", "uint32_t", pentmask4(), [16, 16]))
    print("wrote", OUT, "edges3", len(EDGES3), "edges4", len(EDGES4))


if __name__ == "__main__":
    main()
