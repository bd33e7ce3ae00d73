"""Logic parity on the CPU: the product's per-path device functions (gnx_path.cuh, compiled for the
host by tests/emul) against the UNMODIFIED reference on the same scene and the same sample indices.

Bars (BASELINE.json north_star): primary-hit primitive IDs equal on >= 99.99 % of pixels; image
rel-MSE <= 1e-3.  On the CPU both sides share libm and neither contracts FMAs, so the test is much
tighter: hits must be identical and per-sample radiance equal to float rounding."""
import numpy as np
import pytest

from _harness import grid, rel_mse
from gnxraytracer_b200.api import RenderParams


@pytest.mark.parametrize("preset,res", [("cornell", 64), ("cornell_on", 48), ("dragon", 96), ("dragon_metal", 64),
                                        ("nano", 64), ("nano_thin", 64)])
def test_primary_hits_and_radiance(ref, emul, preset, res):
    rs = ref.scene(preset, res, res, 4)
    es = emul.scene(rs.desc)
    px, py = grid(res, res)
    params = RenderParams.make(res, res, 4, max_depth=5)
    for s in (0, 3):
        sm = np.full(px.size, s, np.int32)
        rgb, prim = rs.reference_samples(px, py, sm, max_depth=5)
        hits = rs.to_original(es.primary_hits(res, res, s))
        assert np.mean(hits == prim) >= 0.9999, "primary-hit parity"
        mine = es.samples(params, px, py, sm)
        scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
        rel = np.abs(mine - rgb).max(axis=1) / scale
        # a handful of paths may take a different discrete decision; none may be systematically off
        assert np.mean(rel < 1e-4) >= 0.999, f"per-sample radiance parity {np.mean(rel < 1e-4)}"
        assert abs(mine.mean() - rgb.mean()) <= 1e-3 * abs(rgb.mean())
    rs.close(); es.close()


def test_whole_image_matches_reference_render(ref, emul):
    """The reference's own Render() (pixel loop + FrameBuffer running mean) against the emulated
    product pipeline, Cornell 48x48 x 8 spp: rel-MSE far below the 1e-3 bar."""
    rs = ref.scene("cornell", 48, 48, 8)
    img_ref, _ = rs.render_reference(max_depth=5)
    es = emul.scene(rs.desc)
    img, st = es.render(RenderParams.make(48, 48, 8, max_depth=5))
    assert rel_mse(img, img_ref) <= 1e-6
    assert st.paths == 48 * 48 * 8 and st.rays_extend >= st.paths
    assert np.all(img[..., 3] == 1)
    rs.close(); es.close()


@pytest.mark.parametrize("depth", [0, 1, 2])
def test_max_depth_edge_cases(ref, emul, depth):
    rs = ref.scene("cornell", 32, 32, 2)
    es = emul.scene(rs.desc)
    px, py = grid(32, 32)
    sm = np.zeros(px.size, np.int32)
    rgb, _ = rs.reference_samples(px, py, sm, max_depth=depth, want_prim=False)
    mine = es.samples(RenderParams.make(32, 32, 2, max_depth=depth), px, py, sm)
    assert np.allclose(mine, rgb, rtol=1e-4, atol=1e-6)
    rs.close(); es.close()


@pytest.mark.parametrize("preset", ["fog", "smoke"])
def test_volpath_matches_reference(ref, emul, preset):
    """VolPathIntegrator with a HomogeneousMedium (Halton) and with the GridDensityMedium of config 4
    (delta / ratio tracking, PCG32 stream sampler): per-sample radiance against VolPathIntegrator::Li."""
    res = 40
    rs = ref.scene(preset, res, res, 4)
    es = emul.scene(rs.desc)
    px, py = grid(res, res)
    p = RenderParams.make(res, res, 4, max_depth=5, integrator=1)
    for s in (0, 3):
        sm = np.full(px.size, s, np.int32)
        rgb, _ = rs.reference_samples(px, py, sm, max_depth=5, want_prim=False)
        mine = es.samples(p, px, py, sm)
        scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
        rel = np.abs(mine - rgb).max(axis=1) / scale
        assert np.mean(rel < 1e-4) >= 0.999
    img_ref, _ = rs.render_reference(max_depth=5)
    img, st = es.render(p)
    assert rel_mse(img, img_ref) <= 1e-6
    assert st.rays_shadow > 0 and st.rays_mis > 0
    rs.close(); es.close()


@pytest.mark.parametrize("preset", ["whitted", "whitted_img", "direct", "direct_area", "direct_all", "direct_all_area"])
def test_whitted_and_direct_lighting_match_reference(ref, emul, preset):
    """WhittedIntegrator / DirectLightingIntegrator(UniformSampleOne) with Point, Spot, Distant and
    SkyBox lights next to the area light; mirror and glass spheres exercise SpecularReflect / SpecularTransmit and the
    depth-first order of the Halton dimensions.  Per-sample radiance against the reference's Li."""
    from _harness import integrator_of
    res = 48
    integ = integrator_of(preset)
    rs = ref.scene(preset, res, res, 4)
    es = emul.scene(rs.desc)
    px, py = grid(res, res)
    p = RenderParams.make(res, res, 4, max_depth=5, integrator=integ)
    for s in (0, 3):
        sm = np.full(px.size, s, np.int32)
        rgb, prim = rs.reference_samples(px, py, sm, max_depth=5)
        hits = rs.to_original(es.primary_hits(res, res, s))
        assert np.mean(hits == prim) >= 0.9999
        mine = es.samples(p, px, py, sm)
        scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
        rel = np.abs(mine - rgb).max(axis=1) / scale
        assert np.mean(rel < 1e-4) >= 0.999, f"per-sample radiance parity {np.mean(rel < 1e-4)}"
        assert abs(mine.mean() - rgb.mean()) <= 1e-3 * abs(rgb.mean())
        assert rgb.mean() > 0.01
    img_ref, _ = rs.render_reference(max_depth=5)
    img, st = es.render(p)
    assert rel_mse(img, img_ref) <= 1e-6
    assert st.rays_shadow > 0
    rs.close(); es.close()


@pytest.mark.parametrize("depth", [1, 2])
def test_whitted_depth_limits(ref, emul, depth):
    rs = ref.scene("whitted", 32, 32, 2)
    es = emul.scene(rs.desc)
    px, py = grid(32, 32)
    sm = np.zeros(px.size, np.int32)
    rgb, _ = rs.reference_samples(px, py, sm, max_depth=depth, want_prim=False)
    mine = es.samples(RenderParams.make(32, 32, 2, max_depth=depth, integrator=2), px, py, sm)
    assert np.allclose(mine, rgb, rtol=1e-4, atol=1e-6)
    rs.close(); es.close()


@pytest.mark.parametrize("preset,w,h,spp", [("cornell", 70, 37, 3), ("dragon", 33, 90, 1), ("whitted", 61, 40, 2)])
def test_non_square_odd_sizes(ref, emul, preset, w, h, spp):
    """Ragged shapes: widths and heights that are not multiples of anything (Halton base scales that do not divide the
    resolution, a sample count below the warp width) through Render() of both sides."""
    from _harness import INTEGRATOR_OF, SCENES
    integ = INTEGRATOR_OF.get(SCENES[preset][0], 0)
    rs = ref.scene(preset, w, h, spp)
    es = emul.scene(rs.desc)
    img_ref, _ = rs.render_reference(max_depth=4)
    img, st = es.render(RenderParams.make(w, h, spp, max_depth=4, integrator=integ))
    assert img.shape == (h, w, 4) and st.paths == w * h * spp
    assert rel_mse(img, img_ref) <= 1e-6
    rs.close(); es.close()


@pytest.mark.parametrize("strategy", [0, 2])
def test_uniform_and_power_light_distributions(ref, emul, strategy):
    """lightSampleStrategy "uniform" and "power" (UniformLightDistribution / PowerLightDistribution,
    core/LightDistribution.cpp:15-50) on a Cornell box with two emitters of unequal Light::Power(): per-sample radiance
    against PathIntegrator::Li built with that strategy; the two strategies must differ from each other."""
    res = 48
    rs = ref.scene("cornell_2l", res, res, 4)
    rs.set_light_strategy(strategy)
    es = emul.scene(rs.desc)
    px, py = grid(res, res)
    sm = np.full(px.size, 1, np.int32)
    rgb, _ = rs.reference_samples(px, py, sm, max_depth=5, want_prim=False)
    mine = es.samples(RenderParams.make(res, res, 4, max_depth=5, light_strategy=strategy), px, py, sm)
    other = es.samples(RenderParams.make(res, res, 4, max_depth=5, light_strategy=2 - strategy), px, py, sm)
    scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
    rel = np.abs(mine - rgb).max(axis=1) / scale
    assert np.mean(rel < 1e-4) >= 0.999, f"per-sample radiance parity {np.mean(rel < 1e-4)}"
    assert np.mean(np.abs(other - rgb).max(axis=1) / scale < 1e-4) < 0.9, "the strategies should sample differently"
    rs.close(); es.close()


@pytest.mark.parametrize("preset,depth", [("ui_path", 15), ("ui_whitted", 5), ("lights_path", 5), ("lights_path_img", 5)])
def test_ui_scene_and_delta_skybox_lights_under_path(ref, emul, preset, depth):
    """The reference UI's live scene (ui/RenderThread.cpp:60-164: mesh inside the Cornell box, area light + SkyBoxLight)
    under the integrator lines :163 (Whitted, maxDepth 5) and :164 (Path, maxDepth 15), and the lights room under the
    wavefront PathIntegrator: EstimateDirect's delta branch (core/Integrator.cpp:148,159) and SkyBoxLight (Pdf_Li = 0, Le
    on escape).  Per-sample radiance against the reference's own Li, image against its Render, all three strategies."""
    from _harness import integrator_of
    res = 40
    integ = integrator_of(preset)
    rs = ref.scene(preset, res, res, 4)
    es = emul.scene(rs.desc)
    px, py = grid(res, res)
    p = RenderParams.make(res, res, 4, max_depth=depth, integrator=integ)
    sm = np.full(px.size, 2, np.int32)
    rgb, prim = rs.reference_samples(px, py, sm, max_depth=depth)
    assert np.mean(rs.to_original(es.primary_hits(res, res, 2)) == prim) >= 0.9999
    if preset.startswith("ui"):
        assert np.all(prim >= 0), "the UI camera looks into the box: every camera ray hits"
    mine = es.samples(p, px, py, sm)
    scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
    assert np.mean(np.abs(mine - rgb).max(axis=1) / scale < 1e-4) >= 0.999
    img_ref, _ = rs.render_reference(max_depth=depth)
    img, st = es.render(p)
    assert rel_mse(img, img_ref) <= 1e-6 and st.rays_shadow > 0
    if integ == 0:
        for strategy in (0, 2):
            rs.set_light_strategy(strategy)
            a, _ = rs.render_reference(max_depth=depth)
            b, _ = emul.scene(rs.desc).render(RenderParams.make(res, res, 4, max_depth=depth, light_strategy=strategy))
            assert rel_mse(b, a) <= 1e-6
    rs.close(); es.close()


def test_scene_kit_ui_scene_equals_the_harness_one(ref, emul):
    """The kit's own "ui" scene (what bench.py's U1 workloads render) against the reference's render of the harness one."""
    from gnxraytracer_b200.api import SceneKit
    res = 40
    rs = ref.scene("ui_path", res, res, 4)
    img_ref, _ = rs.render_reference(max_depth=15)
    sk = SceneKit("ui", res, res, 4, 0, 256, 32)
    img, _ = emul.scene(sk.desc).render(RenderParams.make(res, res, 4, max_depth=15))
    # (another BVH: rays through the room's edges hit either of two walls at the same distance)
    assert rel_mse(img, img_ref) <= 1e-3
    assert np.mean(np.abs(img[..., :3] - img_ref[..., :3]).max(axis=2) < 1e-5) >= 0.99
    rs.close(); sk.close()


@pytest.mark.parametrize("preset", ["whitted_tex", "whitted_tri", "direct_tex", "fog_tex", "fog_tri"])
def test_image_textures_filtered_with_ray_differentials(ref, emul, preset):
    """MIPMap::Lookup with ray differentials (core/MIPMap.h:226-337): EWA and trilinear filtering over the pyramid, fed by
    SurfaceInteraction::ComputeDifferentials (core/Interaction.cpp:65-114) from the camera's offset rays
    (camera/Perspective.cpp:62-112, scaled by 1 / sqrt(spp)) and, under Whitted / DirectLighting, from the offset rays of
    SpecularReflect / SpecularTransmit (core/Integrator.cpp:321-442: mirror and glass spheres with vertex normals in front of
    textured walls).  VolPath keeps the differentials for the camera segment only.  Per-sample radiance against the
    reference's Li, the image against its Render; EWA and trilinear must differ from each other and from plain bilinear."""
    from _harness import integrator_of
    res = 48
    integ = integrator_of(preset)
    rs = ref.scene(preset, res, res, 4)
    es = emul.scene(rs.desc)
    px, py = grid(res, res)
    p = RenderParams.make(res, res, 4, max_depth=5, integrator=integ)
    sm = np.full(px.size, 1, np.int32)
    rgb, _ = rs.reference_samples(px, py, sm, max_depth=5, want_prim=False)
    mine = es.samples(p, px, py, sm)
    scale = np.maximum(np.abs(rgb).max(axis=1), 1e-3)
    assert np.mean(np.abs(mine - rgb).max(axis=1) / scale < 1e-4) >= 0.999
    img_ref, _ = rs.render_reference(max_depth=5)
    img, _ = es.render(p)
    assert rel_mse(img, img_ref) <= 1e-6
    if preset in ("whitted_tex", "fog_tex"):
        other = ref.scene(preset.replace("_tex", "_tri"), res, res, 4)
        img_tri, _ = other.render_reference(max_depth=5)
        assert rel_mse(img, img_tri) > 1e-7, "EWA and trilinear filtering should not coincide"
        other.close()
    rs.close(); es.close()


@pytest.mark.parametrize("preset", ["cornell", "dragon", "whitted", "direct_all", "fog"])
def test_sobol_sampler_matches_the_reference_helpers(ref, emul, preset):
    """GNX_SAMPLER_SOBOL: the Sobol' GlobalSampler built from the reference's SobolIntervalToIndex / SobolSample and its
    generator matrices (samplers/LowDiscrepancy.h:194-252, samplers/SobolMatrices.h:12-17; the class itself is
    gnxraytracer_b200/bridge/SobolSampler.h, running inside the reference's own integrators on the oracle side): sample values
    bit for bit, GetIndexForSample for pixels and samples, and the images of Path / Whitted / UniformSampleAll (sample arrays) /
    VolPath through it."""
    from _harness import integrator_of
    res, spp = 40, 4
    rs = ref.scene(preset, res, res, spp)
    rs.set_sampler(2)
    es = emul.scene(rs.desc)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 1 << 24, 20000).astype(np.int64)
    dim = rng.integers(2, 1024, 20000).astype(np.int32)
    assert np.array_equal(rs.sample_dims(idx, dim).view(np.uint32), es.sample_dims(idx, dim).view(np.uint32))
    for (x, y, s) in ((0, 0, 0), (5, 7, 1), (39, 39, 3), (17, 2, 2)):
        assert rs.sample_index(x, y, s) == es.sample_index(x, y, s)
    img_ref, _ = rs.render_reference(max_depth=5)
    img, _ = es.render(RenderParams.make(res, res, spp, max_depth=5, integrator=integrator_of(preset)))
    assert rel_mse(img, img_ref) <= 1e-6
    rs.set_sampler(0)
    img_halton, _ = rs.render_reference(max_depth=5)
    assert rel_mse(img, img_halton) > 1e-6, "another sampler, another image"
    rs.close(); es.close()
