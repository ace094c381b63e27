"""CPU: ctr_wire_format (host-only entry point of the C ABI) against the reference's own string building,
restated verbatim from html_demo.py:118-161 and morph_geometry.py:91-128 (plain str() joins)."""
import numpy as np
import pytest

from contourist_b200 import html_demo, morph_geometry, wire


class FakeContour(object):
    def __init__(self, points, triangles):
        self.points, self.triangles = points, triangles

    def get_points_and_triangles(self):
        return self.points, self.triangles


def reference_three_json(points, triangles):
    "html_demo.py:133-161, as written there"
    faces = []
    for triangle in triangles:
        faces.append("0")
        for index in triangle:
            faces.append(str(index))
    vertices = []
    for point in points:
        for coordinate in point:
            vertices.append(str(coordinate))
    return html_demo.json_template % {"faces": "[%s]" % (",\n".join(faces)), "vertices": "[%s]" % (",\n".join(vertices))}


def nasty_floats(rng, n):
    mant = rng.standard_normal(n)
    x = np.concatenate([mant * 10.0 ** rng.integers(-40, 40, n), rng.uniform(-2, 2, n), rng.integers(-10 ** 6, 10 ** 6, n).astype(float),
                        np.array([0.0, -0.0, 1e-5, 1e-4, 9.999e-5, 1e15, 1e16, 9999999999999998.0, 123456789012345680.0, 0.1, 1 / 3.0,
                                  5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1e22, 1e23, 100000.0, 0.001])])
    return x


def test_float_repr_is_pythons():
    x = nasty_floats(np.random.default_rng(1), 20000)
    assert wire.format_rows(x, row_sep=";") == "[" + ";".join(str(float(v)) for v in x) + "]"
    with np.errstate(over="ignore"):
        f = x.astype(np.float32)
    f = f[np.isfinite(f)]
    assert wire.format_rows(f, row_sep=";") == "[" + ";".join(str(float(v)) for v in f) + "]"
    assert wire.format_rows(np.array([np.inf, -np.inf, np.nan])) == "[inf,\n-inf,\nnan]"


def test_int_rows_and_thread_counts_agree():
    rng = np.random.default_rng(2)
    a = rng.integers(-2 ** 40, 2 ** 40, size=(200001, 3))
    want = "[%s]" % (",\n".join(",".join(str(y) for y in x) for x in a.tolist()))
    for threads in (1, 3, 0):
        assert wire.format_rows(a, threads=threads) == want
    a32 = rng.integers(-2 ** 31, 2 ** 31 - 1, size=(70000, 2)).astype(np.int32)
    a32[0] = (-2 ** 31, 2 ** 31 - 1)
    assert wire.format_rows(a32) == "[%s]" % (",\n".join(",".join(str(y) for y in x) for x in a32.tolist()))
    assert wire.format_rows(np.zeros((0, 3), np.int32)) == "[]"
    assert wire.format_rows(np.array([[7]], np.uint32)) == "[7]"
    with pytest.raises(ValueError):
        wire.format_rows(np.zeros((2, 2, 2)))
    with pytest.raises(ValueError):
        wire.format_rows(np.array([["a"]]))


def test_emit_three_json_bytes_equal_reference_loops():
    rng = np.random.default_rng(3)
    pts = nasty_floats(rng, 30000)
    pts = pts[:len(pts) // 3 * 3].reshape(-1, 3)
    tris = rng.integers(0, len(pts), size=(70001, 3)).astype(np.int32)
    got = html_demo.emit_three_json(FakeContour(pts, tris))
    assert got == reference_three_json([list(map(float, p)) for p in pts], [tuple(int(i) for i in t) for t in tris])
    # the reference hands lists of numpy rows / index tuples to the same code
    assert html_demo.emit_three_json(FakeContour(list(pts), [tuple(t) for t in tris.tolist()])) == got
    assert html_demo.emit_three_json(FakeContour(np.zeros((0, 3)), np.zeros((0, 3), np.int32))) == reference_three_json([], [])


def test_grid_html_page_lists_equal_reference():
    rng = np.random.default_rng(4)
    pts = rng.standard_normal((1000, 3)) * 10.0 ** rng.integers(-6, 6, (1000, 1))
    tris = rng.integers(0, 1000, size=(2000, 3))
    page = html_demo.grid_html_page(FakeContour(pts, tris))
    # html_demo.py:120-121: ",\n    ".join(map(str, map(list, points))) on rows of Python floats / ints
    vertices = "[%s]" % (",\n    ".join(map(str, ([float(c) for c in p] for p in pts))))
    indices = "[%s]" % (",\n    ".join(map(str, ([int(i) for i in t] for t in tris))))
    assert "var MESH_POINTS = %s;" % vertices in page and "var MESH_FACES = %s;" % indices in page


def test_flatten_json_list_bytes():
    rng = np.random.default_rng(5)
    a = rng.integers(0, 999999, size=(100000, 4))
    assert morph_geometry.flatten_json_list(a) == "[%s]" % (",\n".join(",".join(str(y) for y in x) for x in a.tolist()))


def test_emit_three_json_bytes_equal_the_reference_function():
    """html_demo.emit_three_json against the text the unmodified reference's emit_three_json produced for the final
    meshes of two 3D goldens (sha256 fixtures, tests/golden/make_golden.py json4d)."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    want = json.load(open(os.path.join(GOLDEN, "mt3d_three_json.json")))
    for name, w in want.items():
        g = np.load(os.path.join(GOLDEN, "mt3d_%s.npz" % name))
        text = html_demo.emit_three_json(FakeContour(g["final_points"], g["final_tris"]))
        assert len(text) == w["length"] and hashlib.sha256(text.encode()).hexdigest() == w["sha256"]
