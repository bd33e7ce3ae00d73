"""Parity at the BASELINE configs' STATED sizes (run on the B200 box with `-m gpu`): every config of BASELINE.json at its
own resolution, through the drop-in class gnx::CUDAPathIntegrator on the reference's own pbr::Scene, against the
UNMODIFIED reference (oracle/_ref) running on the box's host cores.

  C1  Cornell 512 x 512 x 16 spp, maxDepth 5: the whole job on both sides (4.19 M paths): rel-MSE, 3 sigma, primary hits.
  C2  dragon-class mesh (872 448 triangles) + MonValley, 1024 x 1024: primary hits over EVERY pixel, image at 4 spp.
  C3  Disney + ImageTexture mesh + TropicalRuins, 1920 x 1080: primary hits over every pixel, image at 2 spp.
  C4  VolPath, grid medium in fog, 1024 x 1024: primary hits over every pixel, image at 2 spp.
  C5  dragon-class mesh, 3840 x 2160: primary hits over every pixel (1 spp).
  U1  the reference UI's live scene (mesh inside the Cornell box, 100 % coverage), 1024 x 1024, Path (maxDepth 15) and
      Whitted (maxDepth 5): primary hits over every pixel, image at 2 spp.

Bars (BASELINE.json north_star): primary-hit primitive IDs equal on >= 99.99 % of pixels, rel-MSE <= 1e-3, per-pixel means
within 3 sigma of the reference estimate.  The reduced sample counts of C2-C4 / U1 are what the CPU finishes in seconds; the
GPU reproduces the reference's sample stream sample by sample, so equal-spp images are compared directly.
"""
import numpy as np
import pytest

from _harness import grid, rel_mse

pytestmark = pytest.mark.gpu


def _hits_agree(rs, width, height, max_depth):
    px, py = grid(width, height)
    _, prim = rs.reference_samples(px, py, np.zeros(px.size, np.int32), max_depth=max_depth, want_rgb=False)
    hits = rs.to_original(rs.cuda_primary_hits(0))
    return float(np.mean(hits == prim)), float(np.mean(prim >= 0))


def test_config1_cornell_in_full(ref):
    res, spp = 512, 16
    rs = ref.scene("cornell_full", res, res, spp)
    img_ref, _ = rs.render_reference(max_depth=5)
    img, _, st = rs.render_cuda(max_depth=5)
    assert st.paths == res * res * spp
    r = rel_mse(img, img_ref)
    assert r <= 1e-3, f"rel-MSE {r}"
    agree, coverage = _hits_agree(rs, res, res, 5)
    assert agree >= 0.9999, f"primary-hit agreement {agree}"
    assert coverage == 1.0  # closed box towards the camera
    px, py = grid(res, res)
    sel = np.random.default_rng(3).choice(px.size, 1024, replace=False)
    samples = np.stack([rs.reference_samples(px[sel], py[sel], np.full(sel.size, s, np.int32), want_prim=False)[0] for s in range(spp)])
    sigma = samples.std(axis=0, ddof=1) / np.sqrt(spp) + 1e-4
    assert np.mean(np.abs(img.reshape(-1, 4)[sel, :3] - img_ref.reshape(-1, 4)[sel, :3]) <= 3 * sigma) >= 0.999
    rs.close()


@pytest.mark.parametrize("preset,width,height,spp,depth,min_cov", [
    ("dragon_full", 1024, 1024, 4, 5, 0.05),        # C2
    ("nano_full", 1920, 1080, 2, 5, 0.05),          # C3
    ("smoke", 1024, 1024, 2, 5, 0.05),              # C4
    ("ui_path_full", 1024, 1024, 2, 15, 1.0),       # U1, PathIntegrator line of ui/RenderThread.cpp:164
    ("ui_whitted_full", 1024, 1024, 2, 5, 1.0),     # U1, the UI's default WhittedIntegrator (:163)
])
def test_configs_at_their_stated_resolution(ref, preset, width, height, spp, depth, min_cov):
    rs = ref.scene(preset, width, height, spp)
    agree, coverage = _hits_agree(rs, width, height, depth)
    assert agree >= 0.9999, f"{preset}: primary-hit agreement {agree} over {width * height} pixels"
    assert coverage >= min_cov
    img_ref, _ = rs.render_reference(max_depth=depth)
    img, _, st = rs.render_cuda(max_depth=depth)
    assert st.paths == width * height * spp
    r = rel_mse(img, img_ref)
    assert r <= 1e-3, f"{preset}: rel-MSE {r}"
    # per-pixel agreement, not only in the mean: equal sample streams give equal pixels up to libm / rounding
    close = np.abs(img[..., :3] - img_ref[..., :3]).max(axis=2) <= 1e-3 * np.maximum(1.0, img_ref[..., :3].max(axis=2))
    assert np.mean(close) >= 0.995, f"{preset}: {np.mean(close)} of the pixels agree"
    rs.close()


def test_config5_primary_hits_at_3840x2160(ref):
    width, height = 3840, 2160
    rs = ref.scene("dragon_full", width, height, 1)
    agree, coverage = _hits_agree(rs, width, height, 5)
    assert agree >= 0.9999, f"primary-hit agreement {agree} over {width * height} pixels"
    assert 0.03 < coverage < 0.5
    rs.close()


@pytest.mark.parametrize("scene,arg0,res,spp,depth", [("dragon", 0, 1024, 2, 5), ("ui", 0, 1024, 1, 15)])
def test_wide_closest_hit_is_bit_equal_at_full_size(scene, arg0, res, spp, depth):
    """GNX_CLOSEST_BVH8=1 (camera and extension rays through the compressed 8-wide tree, flagged rays traced again in
    reference order; gnx_bvh8.cuh) against the default (reference-order two-child tree) on the full-size meshes of C2 and
    U1: every pixel bit-equal.  1024 x 1024 includes the ray through the exact image centre (d = (0, 0, -1)), whose
    infinite 1 / d once made the wide traversal visit every node."""
    import os
    from gnxraytracer_b200.api import Context, RenderParams, SceneKit
    sk = SceneKit(scene, res, res, spp, arg0, 0, 0)
    p = RenderParams.make(res, res, spp, max_depth=depth)
    imgs = []
    for wide in ("0", "1"):
        os.environ["GNX_CLOSEST_BVH8"] = wide
        try:
            c = Context(0)
        finally:
            del os.environ["GNX_CLOSEST_BVH8"]
        c.upload(sk.desc)
        img, st = c.render(p)
        assert st.paths == res * res * spp
        imgs.append(img)
        c.close()
    assert np.array_equal(imgs[0], imgs[1])
    sk.close()
