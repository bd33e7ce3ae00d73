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


def _write_obj_with_materials(tmp_path):
    """A small OBJ + MTL: a matte floor quad (polygon), a plastic box, a mirror and a glass tetrahedron, one face without usemtl."""
    (tmp_path / "scene.mtl").write_text(
        "newmtl floor\nKd 0.6 0.55 0.5\nillum 1\n"
        "newmtl shiny\nKd 0.1 0.3 0.7\nKs 0.5 0.5 0.5\nNs 200\nillum 2\n"
        "newmtl chrome\nKd 0 0 0\nKs 0.9 0.9 0.9\nillum 3\n"
        "newmtl glass\nKd 0 0 0\nKs 1 1 1\nNi 1.5\nd 0.1\nillum 7\n")
    v = ["v -3 -1 -3", "v 3 -1 -3", "v 3 -1 3", "v -3 -1 3"]                      # floor 1-4
    box = [(-1.6, -1, -0.6), (-0.4, -1, -0.6), (-0.4, -1, 0.6), (-1.6, -1, 0.6), (-1.6, 0.2, -0.6), (-0.4, 0.2, -0.6), (-0.4, 0.2, 0.6), (-1.6, 0.2, 0.6)]
    v += ["v %g %g %g" % p for p in box]                                          # 5-12
    for cx in (0.6, 1.9):                                                         # two tetrahedra 13-16, 17-20
        v += ["v %g -1 -0.5" % (cx - 0.5), "v %g -1 -0.5" % (cx + 0.5), "v %g -1 0.6" % cx, "v %g 0.3 0" % cx]
    f = ["usemtl floor", "f 1 2 3 4", "usemtl shiny",
         "f 5 6 7 8", "f 9 12 11 10", "f 5 9 10 6", "f 6 10 11 7", "f 7 11 12 8", "f 8 12 9 5",
         "usemtl chrome", "f 13 14 16", "f 14 15 16", "f 15 13 16", "f 13 15 14",
         "usemtl glass", "f 17 18 20", "f 18 19 20", "f 19 17 20", "f -4 -2 -3"]
    path = tmp_path / "scene.obj"
    path.write_text("mtllib scene.mtl\n" + "\n".join(v) + "\n" + "\n".join(f) + "\n")
    return str(path)


def test_obj_with_mtl_feeds_the_reference_classes_and_the_kit_alike(ref, emul, tmp_path):
    """one OBJ + MTL file, two consumers.  The oracle harness reads it with the kit's reader and builds the
    REFERENCE's objects from it (TriangleMesh / GeometricPrimitive per material; MatteMaterial / PlasticMaterial / MirrorMaterial /
    GlassMaterial from the MTL entries through gnxsk::material_recipe) — the loader the reference lacks; the scene kit turns the
    same recipe into gnx_material records.  Both render the same image."""
    path = _write_obj_with_materials(tmp_path)
    info = mesh_info(path)
    assert info["triangles"] == 2 + 12 + 4 + 4
    res, spp = 64, 4
    rs = ref.scene_named("obj:" + path, res, res, spp)
    sk = SceneKit("obj:" + path, res, res, spp)
    assert sk.num_prims == rs.lib.gnxh_scene_num_prims(rs.h) == info["triangles"]
    img_ref, _ = rs.render_reference(max_depth=5)
    img, _ = emul.scene(sk.desc).render(RenderParams.make(res, res, spp, max_depth=5))
    assert rel_mse(img, img_ref) <= 1e-3
    assert np.mean(np.abs(img[..., :3] - img_ref[..., :3]).max(axis=2) < 1e-4) >= 0.98  # (another BVH: ties along shared edges)
    # the four materials are really there: a render with everything in the default material looks different
    (tmp_path / "plain.obj").write_text("\n".join(l for l in open(path).read().splitlines() if not l.startswith(("usemtl", "mtllib"))) + "\n")
    plain = SceneKit("obj:" + str(tmp_path / "plain.obj"), res, res, spp)
    img_plain, _ = emul.scene(plain.desc).render(RenderParams.make(res, res, spp, max_depth=5))
    assert rel_mse(img_plain, img_ref) > 1e-3
    rs.close(); sk.close(); plain.close()


def test_mtl_errors_are_reported(tmp_path):
    (tmp_path / "a.obj").write_text("mtllib a.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl nope\nf 1 2 3\n")
    (tmp_path / "a.mtl").write_text("newmtl yes\nKd 1 0 0\n")
    with pytest.raises(RuntimeError, match="unknown material"):
        SceneKit("obj:" + str(tmp_path / "a.obj"), 8, 8, 1)
    (tmp_path / "c.mtl").write_text("newmtl m\nKd oops\n")
    (tmp_path / "c.obj").write_text("mtllib c.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    with pytest.raises(RuntimeError, match="malformed 'Kd'"):
        SceneKit("obj:" + str(tmp_path / "c.obj"), 8, 8, 1)
    # a library that is simply not there is tolerated: the faces keep the scene's default material
    (tmp_path / "b.obj").write_text("mtllib missing.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl whatever\nf 1 2 3\n")
    assert SceneKit("obj:" + str(tmp_path / "b.obj"), 8, 8, 1).num_prims == 1


@pytest.mark.gpu
def test_file_loaded_meshes_on_the_device(ref, tmp_path):
    """The ingest paths on the GPU: a .3d file read by the reference's plyInfo on one side and by the kit on the other, and the
    OBJ + MTL scene — rendered by libgnxrt from the kit's description, compared with the reference's render of its own objects."""
    from gnxraytracer_b200.api import Context
    ctx = Context(0)
    res, spp = 128, 8
    path3d = str(tmp_path / "knot.3d")
    write_knot_3d(path3d, 256, 32)
    for name in ("dragon3d:" + path3d, "obj:" + _write_obj_with_materials(tmp_path)):
        rs = ref.scene_named(name, res, res, spp)
        sk = SceneKit(name, res, res, spp)
        ctx.upload(sk.desc)
        img, st = ctx.render(RenderParams.make(res, res, spp, max_depth=5))
        img_ref, _ = rs.render_reference(max_depth=5)
        assert st.paths == res * res * spp
        assert rel_mse(img, img_ref) <= 1e-3, name
        px, py = grid(res, res)
        _, prim = rs.reference_samples(px, py, np.zeros(px.size, np.int32), want_rgb=False)
        hits = ctx.primary_hits(RenderParams.make(res, res, spp), 0).ravel()
        assert np.mean((hits >= 0) == (prim >= 0)) >= 0.9999   # same silhouette (primitive numbering differs between the two loaders' orders)
        # and through the drop-in class on the reference's objects built from the file
        img2, _, _ = rs.render_cuda(max_depth=5)
        assert rel_mse(img2, img_ref) <= 1e-3
        rs.close(); sk.close()
    ctx.close()
