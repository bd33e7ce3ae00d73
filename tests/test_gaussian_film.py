"""Gaussian reconstruction film (north_star stage 5, gnx_film.cuh).

The reference owns filters/GaussianFilter but its Render() box-averages, so the oracle here is the Film::AddSample splat
of the reference's lineage around the reference's OWN camera samples, Li and GaussianFilter::Evaluate
(oracle/ref_harness.cpp::gnxh_reference_gaussian_film).  Tolerances: filter weights <= 4 ulp of the filter's peak (CUDA /
glibc expf); images rel-MSE <= 1e-6 — the sums run in a different order (gather per pixel against splat per sample)."""
import numpy as np
import pytest

from _harness import rel_mse
from gnxraytracer_b200.api import FILM_GAUSSIAN, FILM_GAUSSIAN_SUMS, RenderParams


@pytest.mark.parametrize("radius,alpha", [(2.0, 2.0), (1.5, 0.5), (3.0, 1.0)])
def test_filter_weights_match_the_reference_class(ref, emul, radius, alpha):
    rng = np.random.default_rng(3)
    n = 20000
    x = rng.uniform(-radius, radius, n).astype(np.float32)
    y = rng.uniform(-radius, radius, n).astype(np.float32)
    x[:4] = [0, radius, -radius, 0.5]
    a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
    ref.lib.gnxh_reference_gaussian_eval(radius, alpha, n, x.ctypes.data, y.ctypes.data, a.ctypes.data)
    emul.lib.gnxe_gaussian_eval(radius, alpha, n, x.ctypes.data, y.ctypes.data, b.ctypes.data)
    assert np.array_equal(a, b), "same libm on the CPU: bit-identical"
    assert a[0] > 0 and a[1] == 0 and a[2] == 0


@pytest.mark.parametrize("preset,w,h,spp,radius,alpha,integ", [("cornell", 48, 40, 4, 2.0, 2.0, 0), ("dragon", 37, 45, 3, 1.5, 0.5, 0),
                                                              ("fog", 32, 32, 2, 2.0, 2.0, 1), ("whitted", 40, 32, 2, 2.5, 1.0, 2)])
def test_gaussian_film_matches_reference_splat(ref, emul, preset, w, h, spp, radius, alpha, integ):
    rs = ref.scene(preset, w, h, spp)
    img_ref, sums_ref = rs.reference_gaussian_film(radius, alpha, max_depth=4)
    es = emul.scene(rs.desc)
    p = RenderParams.make(w, h, spp, max_depth=4, integrator=integ, film=FILM_GAUSSIAN, filter_radius=radius, filter_alpha=alpha)
    img, _ = es.render(p)
    assert rel_mse(img, img_ref) <= 1e-6
    assert np.allclose(img[..., :3], img_ref[..., :3], rtol=2e-4, atol=1e-5)
    assert np.all(img[..., 3] == 1)
    p.film = FILM_GAUSSIAN_SUMS
    sums, _ = es.render(p)
    assert np.allclose(sums[..., 3], sums_ref[..., 3], rtol=1e-5), "filter weight sums"
    # it is a different image from the box average
    box, _ = es.render(RenderParams.make(w, h, spp, max_depth=4, integrator=integ))
    assert rel_mse(box, img_ref) > 1e-4
    rs.close(); es.close()


# ---- on the B200, through the C ABI and through the drop-in class ------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("preset,w,h,spp,radius,alpha,integ", [("cornell", 96, 80, 8, 2.0, 2.0, 0), ("dragon", 101, 67, 5, 1.5, 0.5, 0),
                                                              ("fog", 48, 48, 4, 2.0, 2.0, 1), ("whitted", 64, 48, 4, 2.5, 1.0, 2),
                                                              ("cornell", 50, 30, 2, 6.0, 0.1, 0), ("cornell", 33, 20, 9, 1.0, 3.0, 0)])
def test_gpu_gaussian_film_matches_reference_splat(ref, preset, w, h, spp, radius, alpha, integ):
    from gnxraytracer_b200.api import Context
    rs = ref.scene(preset, w, h, spp)
    img_ref, sums_ref = rs.reference_gaussian_film(radius, alpha, max_depth=4)
    ctx = Context(0)
    ctx.upload(rs.desc)
    p = RenderParams.make(w, h, spp, max_depth=4, integrator=integ, film=FILM_GAUSSIAN, filter_radius=radius, filter_alpha=alpha)
    img, st = ctx.render(p)
    assert st.paths == w * h * spp
    assert rel_mse(img, img_ref) <= 1e-6
    assert np.all(img[..., 3] == 1)
    # sample batches: the sums of the batches add up to the one-batch sums (float rounding of the partial sums only)
    p1 = RenderParams.make(w, h, spp, max_depth=4, integrator=integ, film=FILM_GAUSSIAN, filter_radius=radius, filter_alpha=alpha,
                           batch_spp=1)
    img1, _ = ctx.render(p1)
    assert np.allclose(img1, img, rtol=1e-4, atol=1e-6)
    again, _ = ctx.render(p)
    assert np.array_equal(again, img), "deterministic: gather in a fixed order, no float atomics"
    p.film = FILM_GAUSSIAN_SUMS
    sums, _ = ctx.render(p)
    assert np.allclose(sums[..., 3], sums_ref[..., 3], rtol=1e-5)
    assert np.allclose(sums[..., :3], sums_ref[..., :3], rtol=2e-4, atol=1e-5)
    # two sample ranges (two ranks of an N-GPU job): their sums add up to the whole job's
    if spp % 2 == 0:
        halves = []
        for first in (0, spp // 2):
            q = RenderParams.make(w, h, spp // 2, first_sample=first, max_depth=4, integrator=integ, film=FILM_GAUSSIAN_SUMS,
                                  filter_radius=radius, filter_alpha=alpha)
            halves.append(ctx.render(q)[0])
        assert np.allclose(halves[0] + halves[1], sums, rtol=1e-4, atol=1e-6)
    ctx.close()
    # the drop-in class with SetGaussianFilter
    rs.set_gaussian_filter(radius, alpha)
    img2, _, _ = rs.render_cuda(max_depth=4)
    assert np.array_equal(img2[..., :3], img[..., :3])  # (the FrameBuffer sink only receives r, g, b)
    rs.close()


@pytest.mark.gpu
def test_gpu_gaussian_film_rejects_bad_filters(ref):
    from gnxraytracer_b200.api import Context, GnxError
    rs = ref.scene("cornell", 32, 32, 1)
    ctx = Context(0)
    ctx.upload(rs.desc)
    for radius, alpha in ((0.0, 1.0), (-1.0, 1.0), (100.0, 1.0), (2.0, -1.0)):
        with pytest.raises(GnxError):
            ctx.render(RenderParams.make(32, 32, 1, film=FILM_GAUSSIAN, filter_radius=radius, filter_alpha=alpha))
    ctx.close(); rs.close()


@pytest.mark.gpu
def test_gpu_gaussian_film_full_size_properties(monkeypatch):
    """Config 2 at full size (872 448 triangles, 1024 x 1024): the tiled shared-memory gather against the plain per-pixel
    gather (the function the CPU emulation runs, GNX_FILM_SIMPLE=1), additivity of the unresolved sums over sample ranges,
    run-to-run determinism, and the filter leaves a constant region constant (the environment far from the mesh is smooth:
    filtered and box images agree there to the noise level)."""
    from gnxraytracer_b200.api import Context, SceneKit
    res, spp = 1024, 4
    sk = SceneKit("dragon", res, res, spp)
    mk = lambda film, first=0, n=spp: RenderParams.make(res, res, n, first_sample=first, film=film, filter_radius=2.0, filter_alpha=2.0)
    ctx = Context(0)
    ctx.upload(sk.desc)
    sums, st = ctx.render(mk(FILM_GAUSSIAN_SUMS))
    assert st.paths == res * res * spp
    again, _ = ctx.render(mk(FILM_GAUSSIAN_SUMS))
    assert np.array_equal(again, sums)
    a, _ = ctx.render(mk(FILM_GAUSSIAN_SUMS, 0, 2))
    b, _ = ctx.render(mk(FILM_GAUSSIAN_SUMS, 2, 2))
    assert np.allclose(a + b, sums, rtol=1e-4, atol=1e-5)
    img, _ = ctx.render(mk(FILM_GAUSSIAN))
    w = sums[..., 3:4]
    assert np.all(w > 0)
    assert np.allclose(img[..., :3], np.maximum(sums[..., :3] / w, 0), rtol=1e-5, atol=1e-6)
    # interior pixels see the samples of 25 pixels: their weight sum is close to spp x the filter's integral over the plane
    interior = sums[8:-8, 8:-8, 3]
    assert interior.std() / interior.mean() < 0.2
    ctx.close()
    monkeypatch.setenv("GNX_FILM_SIMPLE", "1")
    ctx2 = Context(0)
    ctx2.upload(sk.desc)
    simple, _ = ctx2.render(mk(FILM_GAUSSIAN_SUMS))
    ctx2.close()
    assert np.allclose(simple, sums, rtol=2e-4, atol=1e-5), "tiled and per-pixel gathers"
    sk.close()
