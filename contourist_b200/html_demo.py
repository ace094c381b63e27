"""three.js emitters -- drop-in for contourist/html_demo.py:118-161 (grid_html_page, emit_three_json).

The mesh comes from get_points_and_triangles() of the CUDA engine; the two big number lists are formatted by the
library's host threads (`ctr_wire_format` via wire.format_rows: the same bytes as the reference's str() joins).
"""
import numpy as np

from . import wire

load_three = '<script src="https://cdnjs.cloudflare.com/ajax/libs/three.js/r70/three.min.js"></script>'

three_html_fullscreen = """<!DOCTYPE html>
<html>
<head>
    <title>%(title)s</title>
    %(load_three)s
    <style> body { margin: 0; overflow: hidden; } </style>
</head>
<body>
<div id="%(target_div)s"></div>
<script type="text/javascript">
    function init() {
        var scene = new THREE.Scene();
        var camera = new THREE.PerspectiveCamera(45, window.innerWidth / window.innerHeight, 0.1, 1000);
        var webGLRenderer = new THREE.WebGLRenderer();
        webGLRenderer.setClearColor(new THREE.Color(0xEEEEEE, 1.0));
        webGLRenderer.setSize(window.innerWidth, window.innerHeight);
        var triangulation = make_triangulation();
        scene.add(triangulation);
        camera.position.set(%(camera_x)s, %(camera_y)s, %(camera_z)s);
        camera.lookAt(new THREE.Vector3(0, 0, 0));
        document.getElementById("%(target_div)s").appendChild(webGLRenderer.domElement);
        var step = 0;
        function render() {
            triangulation.rotation.y = step += 0.01;
            requestAnimationFrame(render);
            webGLRenderer.render(scene, camera);
        };
        render();
    };
    window.onload = init;

    function make_triangulation() {
        var vertices = %(vertices)s;
        var indices = %(indices)s;
        var geom = new THREE.Geometry();
        for (var i=0; i<vertices.length; i++) {
            var v = vertices[i];
            geom.vertices.push(new THREE.Vector3(v[0], v[1], v[2]));
        }
        for (var i=0; i<indices.length; i++) {
            var f = indices[i];
            geom.faces.push(new THREE.Face3(f[0], f[1], f[2]));
        }
        geom.computeFaceNormals();
        geom.computeVertexNormals();
        var meshMaterial = new THREE.MeshNormalMaterial();
        meshMaterial.side = THREE.DoubleSide;
        var wireFrameMat = new THREE.MeshBasicMaterial();
        wireFrameMat.wireframe = true;
        return THREE.SceneUtils.createMultiMaterialObject(geom, [meshMaterial, wireFrameMat]);
    };
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
