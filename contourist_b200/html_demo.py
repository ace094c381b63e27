"""three.js emitters -- drop-in for contourist/html_demo.py:118-161 (grid_html_page, emit_three_json).

The mesh comes from get_points_and_triangles() of the CUDA engine; the two big number lists are formatted by the
library's host threads (`ctr_wire_format` via wire.format_rows: the same bytes as the reference's str() joins).
"""
import numpy as np

from . import wire

load_three = '<script src="https://cdnjs.cloudflare.com/ajax/libs/three.js/r70/three.min.js"></script>'

# A stand-alone viewer page for one indexed mesh.  Not a wire format: only the placeholders and the two number lists
# (`vertices` = [[x, y, z], ...], `indices` = [[i, j, k], ...], formatted like Python's str(list)) follow
# html_demo.py:118-131; the page itself is this package's own (flat-shaded BufferGeometry, orbit by dragging).
three_html_fullscreen = """<!DOCTYPE html>
<html>
<head>
<meta charset="utf-8">
<title>%(title)s</title>
%(load_three)s
<style>html, body { height: 100%%; margin: 0; background: #f2f2f2; } #%(target_div)s { position: fixed; inset: 0; }</style>
</head>
<body>
<div id="%(target_div)s"></div>
<script>
(function () {
  var MESH_POINTS = %(vertices)s;
  var MESH_FACES = %(indices)s;
  var host = document.getElementById("%(target_div)s");
  var renderer = new THREE.WebGLRenderer({antialias: true});
  renderer.setClearColor(0xf2f2f2, 1);
  host.appendChild(renderer.domElement);
  var eye = new THREE.PerspectiveCamera(40, 1, 0.05, 5000);
  var world = new THREE.Scene();

  // un-indexed triangle soup: one normal per face, so facets stay visible
  var nf = MESH_FACES.length, xyz = new Float32Array(nf * 9), nrm = new Float32Array(nf * 9);
  var lo = [Infinity, Infinity, Infinity], hi = [-Infinity, -Infinity, -Infinity];
  for (var f = 0; f < nf; f++) {
    var P = [MESH_POINTS[MESH_FACES[f][0]], MESH_POINTS[MESH_FACES[f][1]], MESH_POINTS[MESH_FACES[f][2]]];
    var ux = P[1][0] - P[0][0], uy = P[1][1] - P[0][1], uz = P[1][2] - P[0][2];
    var vx = P[2][0] - P[0][0], vy = P[2][1] - P[0][1], vz = P[2][2] - P[0][2];
    var n = [uy * vz - uz * vy, uz * vx - ux * vz, ux * vy - uy * vx];
    var len = Math.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) || 1;
    for (var c = 0; c < 3; c++) {
      for (var a = 0; a < 3; a++) {
        xyz[f * 9 + c * 3 + a] = P[c][a];
        nrm[f * 9 + c * 3 + a] = n[a] / len;
        lo[a] = Math.min(lo[a], P[c][a]);
        hi[a] = Math.max(hi[a], P[c][a]);
      }
    }
  }
  var soup = new THREE.BufferGeometry();
  soup.addAttribute("position", new THREE.BufferAttribute(xyz, 3));
  soup.addAttribute("normal", new THREE.BufferAttribute(nrm, 3));
  var skin = new THREE.Mesh(soup, new THREE.MeshNormalMaterial({side: THREE.DoubleSide}));
  var pivot = new THREE.Object3D();
  pivot.add(skin);
  if (nf) skin.position.set(-(lo[0] + hi[0]) / 2, -(lo[1] + hi[1]) / 2, -(lo[2] + hi[2]) / 2);
  world.add(pivot);
  eye.position.set(%(camera_x)s, %(camera_y)s, %(camera_z)s);
  eye.lookAt(new THREE.Vector3(0, 0, 0));

  function fit() {
    renderer.setSize(window.innerWidth, window.innerHeight);
    eye.aspect = window.innerWidth / Math.max(window.innerHeight, 1);
    eye.updateProjectionMatrix();
  }
  window.addEventListener("resize", fit);
  fit();

  var drag = null, spin = 0.004;
  host.addEventListener("mousedown", function (e) { drag = [e.clientX, e.clientY]; spin = 0; });
  window.addEventListener("mouseup", function () { drag = null; });
  window.addEventListener("mousemove", function (e) {
    if (!drag) return;
    pivot.rotation.y += (e.clientX - drag[0]) * 0.01;
    pivot.rotation.x += (e.clientY - drag[1]) * 0.01;
    drag = [e.clientX, e.clientY];
  });
  (function frame() {
    pivot.rotation.y += spin;
    renderer.render(world, eye);
    requestAnimationFrame(frame);
  })();
})();
</script>
</body>
</html>
"""

json_template = """
{
    "metadata": {
        "version": 3,
        "type": "Geometry",
        "generator": "GeometryExporter"
    },
    "faces": %(faces)s,
    "vertices": %(vertices)s,
    "normals": [],
    "uvs": []
}
"""


def grid_html_page(gridcontour, title="3d contour", load_three=load_three, x=-30, y=40, z=50):
    (points, triangles) = gridcontour.get_points_and_triangles()
    D = {"title": title, "target_div": "THREE_OUTPUT", "load_three": load_three,
         "camera_x": x, "camera_y": y, "camera_z": z}
    rows = dict(row_prefix="[", col_sep=", ", row_suffix="]", row_sep=",\n    ")       # str(list(row)) per row
    D["vertices"] = wire.format_rows(np.asarray(points, dtype=np.float64).reshape(-1, 3), **rows)
    D["indices"] = wire.format_rows(np.asarray(triangles, dtype=np.int64).reshape(-1, 3), **rows)
    return three_html_fullscreen % D


def emit_three_json(grid_contour):
    "THREE Geometry format 3: faces = [0, i, j, k, ...], vertices flat (html_demo.py:133-161)."
    (points, triangles) = grid_contour.get_points_and_triangles()
    faces = wire.format_rows(np.asarray(triangles, dtype=np.int64).reshape(-1, 3), row_prefix="0,\n", col_sep=",\n")
    vertices = wire.format_rows(np.asarray(points, dtype=np.float64).reshape(-1, 1))
    return json_template % {"faces": faces, "vertices": vertices}
