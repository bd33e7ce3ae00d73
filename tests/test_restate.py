"""Pins the independent CPU restatement (oracle/restate/gnx_restate.cpp): against the unmodified
reference where oracle/_ref was built, and against the committed golden vectors everywhere."""
import os

import numpy as np
import pytest

import _harness
from _harness import grid, rel_mse
from gnxraytracer_b200.api import RenderParams, SceneKit
from gnxraytracer_b200.build import build_oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def restate():
    if not os.path.exists(_harness.RESTATE_LIB):
        build_oracle()
    return _harness.Restate()


def test_restate_halton_matches_golden_bit_exact(restate):
    g = np.load(os.path.join(G, "halton_values.npz"))
    sk = SceneKit("cornell", 96, 96, 4, 0, 2, 0)
    rs = restate.scene(sk.desc)
    v = rs.sample_dims(g["index"], g["dim"])
    assert np.array_equal(v.view(np.uint32), g["value"].view(np.uint32))
    for x, y, s, want in g["pixel_index"]:
        assert rs.sample_index(x, y, s) == want
    rs.close(); sk.close()


@pytest.mark.parametrize("name,args", [("cornell", (0, 2, 0)), ("dragon", (0, 256, 32))])
def test_restate_matches_golden_images(restate, name, args):
    g = np.load(os.path.join(G, f"{name}_96x96_4spp.npz"))
    sk = SceneKit(name, 96, 96, 4, *args)
    rs = restate.scene(sk.desc)
    assert np.mean(rs.primary_hits(96, 96, 0) == g["primary_hit"]) >= 0.999  # scene-kit camera: ulp-level ray differences
    img, counts = rs.render(RenderParams.make(96, 96, 4, max_depth=5))
    assert rel_mse(img, g["image"]) <= 1e-3
    assert counts["rays_extend"] >= 96 * 96 * 4
    rs.close(); sk.close()


@pytest.mark.parametrize("preset,res", [("cornell", 64), ("cornell_on", 48), ("dragon", 96), ("dragon_metal", 64)])
def test_restate_matches_live_reference(ref, restate, preset, res):
    r = ref.scene(preset, res, res, 4)
    rs = restate.scene(r.desc)
    px, py = grid(res, res)
    p = RenderParams.make(res, res, 4, max_depth=5)
    for s in (0, 2):
        sm = np.full(px.size, s, np.int32)
        rgb, prim = r.reference_samples(px, py, sm)
        assert np.array_equal(r.to_original(rs.primary_hits(res, res, s)), prim)
        mine = rs.samples(p, px, py, sm)
        scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
        rel = np.abs(mine - rgb).max(axis=1) / scale
        assert np.mean(rel < 1e-5) >= 0.999, np.mean(rel < 1e-5)
    r.close(); rs.close()


@pytest.mark.parametrize("strategy", [0, 2])
def test_restate_uniform_and_power_light_distributions(ref, restate, strategy):
    """lightSampleStrategy "uniform" / "power" on two emitters of unequal power, restatement against PathIntegrator::Li."""
    res = 40
    r = ref.scene("cornell_2l", res, res, 4)
    r.set_light_strategy(strategy)
    rs = restate.scene(r.desc)
    px, py = grid(res, res)
    sm = np.full(px.size, 2, np.int32)
    rgb, _ = r.reference_samples(px, py, sm, want_prim=False)
    mine = rs.samples(RenderParams.make(res, res, 4, max_depth=5, light_strategy=strategy), px, py, sm)
    scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
    assert np.mean(np.abs(mine - rgb).max(axis=1) / scale < 1e-5) >= 0.999
    r.close(); rs.close()


def test_restate_and_emulated_kernels_agree_on_counters(ref, restate, emul):
    """BVH nodes visited / triangles tested by the kernels' traversal must agree with the
    reference-order traversal to < 1 %."""
    r = ref.scene("cornell", 48, 48, 4)
    p = RenderParams.make(48, 48, 4, max_depth=5)
    a, ca = restate.scene(r.desc).render(p)
    # counters of the reference-order traversal: every query on the two-child tree.  The default build sends the any-hit
    # queries through the compressed 8-wide tree (gnx_bvh8.cuh), GNX_CLOSEST_BVH8=1 the closest-hit queries as well (rays
    # whose best hit has a rival within the tie band are traced again in reference order): the very same image each time
    def emul_with(**env):
        os.environ.update(env)
        try:
            return emul.scene(r.desc)
        finally:
            for k in env:
                del os.environ[k]
    b8, sb8 = emul.scene(r.desc).render(p)
    bc, sbc = emul_with(GNX_CLOSEST_BVH8="1").render(p)
    b, sb = emul_with(GNX_ANYHIT_BVH8="0", GNX_CLOSEST_BVH8="0").render(p)
    assert np.array_equal(b8, b) and np.array_equal(bc, b)
    assert sb8.rays_extend == sb.rays_extend == sbc.rays_extend and sb8.rays_shadow == sb.rays_shadow == sbc.rays_shadow
    assert rel_mse(b, a) <= 1e-6
    assert ca["rays_extend"] == sb.rays_extend and ca["rays_shadow"] == sb.rays_shadow
    assert ca["tris_tested"] == sb.tris_tested
    # Slab tests: the 4-wide layout tests the boxes of a node's grandchildren directly, so it never tests a
    # child's own box: fewer tests than the reference order (which tests every popped node), never more, and
    # the very same leaves in the very same order (triangle counts above are identical).
    assert 0.6 * ca["nodes_visited"] <= int(sb.nodes_visited) <= ca["nodes_visited"]
    r.close()


def test_wide_tree_constrains_rays_parallel_to_an_axis(emul):
    """The camera ray through the exact image centre has d = (0, 0, -1): 1 / d is infinite on two axes.  The compressed
    8-wide tree's slab test once turned that into NaN = "no constraint" and the ray visited every node (gnx_bvh8.cuh caps
    |1 / d|).  A 2 x 2 image's pixel (1, 1), sample 0, is that ray: the wide traversal must not visit more node bytes than
    the two-child tree does, and must find the same hits."""
    sk = SceneKit("dragon", 2, 2, 1, 0, 96, 24)
    p = RenderParams.make(2, 2, 1, max_depth=0)
    res = {}
    for wide in ("0", "1"):
        os.environ["GNX_CLOSEST_BVH8"] = wide
        try:
            es = emul.scene(sk.desc)
        finally:
            del os.environ["GNX_CLOSEST_BVH8"]
        img, st = es.render(p)
        res[wide] = (img, int(st.nodes_visited), es.primary_hits(2, 2, 0))
    assert np.array_equal(res["0"][0], res["1"][0])
    assert np.array_equal(res["0"][2], res["1"][2])
    assert res["1"][1] <= 1.5 * res["0"][1] + 64, (res["0"][1], res["1"][1])
    sk.close()
