"""CPU: the legacy "sequence of morphing triangularizations" (pentatopes.py:370-444, morph_geometry.py:130-330; loaded by
misc/morph_sequence.js) restated in contourist_b200/morph_geometry.py, against the unmodified reference's
iterate_morph_geometry / json_data on a 4D golden (tests/golden/make_golden.py seq4d).  Vertex and triangle ORDER follow
CPython dict / set iteration in the reference; intervals, scales, and the triangles as sets of (start, end) integer
corner positions must be identical."""
import json
import os

import numpy as np

from conftest import GOLDEN
from contourist_b200 import morph_geometry


def as_sets(start, end, tris):
    return set(frozenset((tuple(start[i]), tuple(end[i])) for i in t) for t in tris)


def test_sequence_equals_reference():
    g = np.load(os.path.join(GOLDEN, "seq4d_morph7.npz"))
    morphs = list(morph_geometry.morph_sequence(g["pos"], g["tets"]))
    assert len(morphs) == len(g["minmax"]) == 63
    n_tri = 0
    flips = 0
    for k, m in enumerate(morphs):
        assert (m.min_value, m.max_value) == tuple(g["minmax"][k])
        D = m.json_data()
        assert np.allclose(D["scale"], g["scale"][k], rtol=1e-12, atol=0) and np.allclose(D["shift"], g["shift"][k], rtol=1e-12, atol=0)
        v0, v1 = g["vert_off"][k], g["vert_off"][k + 1]
        t0, t1 = g["tri_off"][k], g["tri_off"][k + 1]
        ref = as_sets(g["start"][v0:v1].tolist(), g["end"][v0:v1].tolist(), g["tris"][t0:t1].tolist())
        mine = as_sets(D["start_positions"], D["end_positions"], D["triangles"])
        assert len(D["triangles"]) == t1 - t0 and len(D["start_positions"]) == v1 - v0
        assert mine == ref
        n_tri += len(D["triangles"])
        # winding: the reference orients each slice outwards (surface_geometry.py:52-140); so does the restatement
        S = np.array(D["start_positions"], dtype=float)
        E = np.array(D["end_positions"], dtype=float)
        M = 0.5 * (S + E)

        def signed(P, T):
            out = {}
            for t in T:
                a, b, c = (P[i] for i in t)
                out[frozenset(tuple(P[i]) for i in t)] = np.sign(np.dot(np.cross(b - a, c - a), (a + b + c) / 3 - P.mean(axis=0)))
            return out
        Mr = 0.5 * (g["start"][v0:v1].astype(float) + g["end"][v0:v1].astype(float))
        a, b = signed(M, D["triangles"]), signed(Mr, g["tris"][t0:t1].tolist())
        common = [x for x in a if x in b and a[x] != 0]
        flips += sum(1 for x in common if a[x] != b[x])
    assert n_tri == int(g["tri_off"][-1]) == 51539
    assert flips <= 0.01 * n_tri                       # same outward rule; the odd triangle where the reference's DFS order decides


def test_sequence_json_text():
    g = np.load(os.path.join(GOLDEN, "seq4d_morph7.npz"))
    morphs = list(morph_geometry.morph_sequence(g["pos"], g["tets"]))
    text = morph_geometry.morph_sequence_json(morphs)
    d = json.loads(text)
    assert d["description"] == "Sequence of morphing triangularizations." and d["number_of_morphs"] == 63
    assert d["min_value"] == 0.0 and d["max_value"] == morphs[-1].max_value
    m0 = d["morph_descriptions"][0]
    assert sorted(m0) == ["description", "end_positions", "max_value", "min_value", "scale", "shift", "start_positions", "triangles"]
    assert len(m0["triangles"]) == 3 * int(g["tri_off"][1]) and len(m0["start_positions"]) == 3 * int(g["vert_off"][1])
    assert max(m0["start_positions"]) <= 9999 and min(m0["start_positions"]) >= 0
