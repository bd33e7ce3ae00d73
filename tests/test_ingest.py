"""Mesh ingestion of the scene kit: the `.3d` reader against the reference's own plyInfo
(shape/plyRead.h:19-48) on the same file, its error handling, and the Wavefront OBJ reader."""
import numpy as np
import pytest

from _harness import grid, rel_mse
from gnxraytracer_b200.api import RenderParams, SceneKit, mesh_info, write_knot_3d


def test_3d_file_through_plyinfo_and_through_the_kit_reader(ref, emul, tmp_path):
    """One .3d file, two readers: the harness hands it to the reference's plyInfo + TriangleMesh exactly as
    ui/ModelList.cpp:49-69 does (x20, Translate(0, -2.9, 0)); the scene kit reads it with its own parser.  Same triangle
    count, same primary hits, same image."""
    path = str(tmp_path / "knot.3d")
    write_knot_3d(path, 96, 24)
    info = mesh_info(path)
    assert info == {"vertices": 96 * 24, "triangles": 2 * 96 * 24, "has_uv": False, "has_normals": False}
    res, spp = 48, 4
    rs = ref.scene_named("dragon3d:" + path, res, res, spp)
    sk = SceneKit("dragon3d:" + path, res, res, spp)
    assert sk.num_prims == rs.lib.gnxh_scene_num_prims(rs.h) == info["triangles"]
    es = emul.scene(sk.desc)
    px, py = grid(res, res)
    _, prim = rs.reference_samples(px, py, np.zeros(px.size, np.int32), want_rgb=False)
    assert np.mean(es.primary_hits(res, res, 0) == prim) >= 0.999
    img_ref, _ = rs.render_reference(max_depth=5)
    img, _ = es.render(RenderParams.make(res, res, spp, max_depth=5))
    assert rel_mse(img, img_ref) <= 1e-3
    # and it is the same scene as the procedural "dragon" with that tessellation (the file holds the knot's vertices)
    sk2 = SceneKit("dragon", res, res, spp, 0, 96, 24)
    img2, _ = emul.scene(sk2.desc).render(RenderParams.make(res, res, spp, max_depth=5))
    assert np.array_equal(img2, img)
    rs.close(); es.close(); sk.close(); sk2.close()


@pytest.mark.parametrize("text,why", [
    ("vertex 3 face 1\n0 0 0 1 0 0 0 1 0\n3 0 1 5\n", "indexes vertex"),          # index out of range
    ("vertex 3 face 1\n0 0 0 1 0 0 0 1\n", "malformed"),                           # short vertex list
    ("ply\nformat ascii 1.0\n", "header"),                                        # a real PLY header is not the .3d layout
    ("vertex 3 face 2\n0 0 0 1 0 0 0 1 0\n3 0 1 2\n", "truncated"),               # fewer faces than announced
    ("", "truncated"),
])
def test_3d_reader_fails_loudly(tmp_path, text, why):
    path = tmp_path / "bad.3d"
    path.write_text(text)
    with pytest.raises(RuntimeError, match=why):
        mesh_info(path)
    with pytest.raises(RuntimeError):
        SceneKit("dragon3d:" + str(path), 8, 8, 1)
    with pytest.raises(RuntimeError, match="cannot open"):
        mesh_info(tmp_path / "missing.3d")


def test_3d_header_order_and_face_token(tmp_path):
    """plyInfo accepts "face" before "vertex" and ignores the token in front of the three indices."""
    path = tmp_path / "tri.3d"
    path.write_text("face 1 vertex 3\n0 0 0  1 0 0  0 1 0\nwhatever 2 1 0\n")
    assert mesh_info(path)["triangles"] == 1 and mesh_info(path)["vertices"] == 3


OBJ = """# a unit quad, a polygon with relative indices, and a triangle
mtllib ignored.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
f 1/1/1 2/2/1 3/3/1 4/4/1
v 0 0 1
v 1 0 1
v 0.5 1 1
f -3/1/1 -2/2/1 -1/3/1
"""


def test_obj_reader(tmp_path, emul):
    path = tmp_path / "quad.obj"
    path.write_text(OBJ)
    info = mesh_info(path)
    assert info == {"vertices": 7, "triangles": 3, "has_uv": True, "has_normals": True}
    (tmp_path / "mixed.obj").write_text(OBJ + "f 1 2 3\n")  # one face without vt / vn: attributes dropped for all
    mixed = mesh_info(tmp_path / "mixed.obj")
    assert mixed["triangles"] == 4 and not mixed["has_uv"] and not mixed["has_normals"]
    (tmp_path / "bad.obj").write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(RuntimeError, match="out of range"):
        mesh_info(tmp_path / "bad.obj")
    sk = SceneKit("obj:" + str(path), 32, 32, 2)
    assert sk.num_prims == 3
    es = emul.scene(sk.desc)
    hits = es.primary_hits(32, 32, 0)
    assert (hits >= 0).sum() > 50, "the fitted mesh is in view"
    img, _ = es.render(RenderParams.make(32, 32, 2, max_depth=3))
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
    es.close(); sk.close()
