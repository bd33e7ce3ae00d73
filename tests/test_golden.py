"""Committed golden vectors (tests/golden/*.npz, generated from the unmodified reference by
tests/golden/make_golden.py) against the product's device functions compiled for the host, on scenes
built by the product's own scene kit.  Needs neither /root/reference nor the oracle build."""
import os

import numpy as np

from _harness import rel_mse
from gnxraytracer_b200.api import RenderParams, SceneKit

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_halton_golden_values_bit_exact(emul):
    g = np.load(os.path.join(G, "halton_values.npz"))
    sk = SceneKit("cornell", 96, 96, 4, 0, 2, 0)
    es = emul.scene(sk.desc)
    v = es.sample_dims(g["index"], g["dim"])
    assert np.array_equal(v.view(np.uint32), g["value"].view(np.uint32))
    for x, y, s, want in g["pixel_index"]:
        assert es.sample_index(x, y, s) == want
    es.close(); sk.close()


def test_cornell_golden_image_and_hits(emul):
    g = np.load(os.path.join(G, "cornell_96x96_4spp.npz"))
    sk = SceneKit("cornell", 96, 96, 4, 0, 2, 0)
    es = emul.scene(sk.desc)
    hits = es.primary_hits(96, 96, 0)
    assert np.mean(hits == g["primary_hit"]) >= 0.999  # scene-kit camera matrices differ by ulps from the reference's
    img, _ = es.render(RenderParams.make(96, 96, 4, max_depth=int(g["max_depth"])))
    assert rel_mse(img, g["image"]) <= 1e-3
    es.close(); sk.close()


def test_dragon_golden_image_and_hits(emul):
    g = np.load(os.path.join(G, "dragon_96x96_4spp.npz"))
    sk = SceneKit("dragon", 96, 96, 4, 0, 256, 32)
    es = emul.scene(sk.desc)
    hits = es.primary_hits(96, 96, 0)
    assert np.mean(hits == g["primary_hit"]) >= 0.999  # own camera matrices: ulp-level ray differences
    img, _ = es.render(RenderParams.make(96, 96, 4, max_depth=5))
    assert rel_mse(img, g["image"]) <= 1e-3
    es.close(); sk.close()


def test_whitted_and_direct_golden_images(emul):
    """the committed reference renders of WhittedIntegrator / DirectLightingIntegrator on the lights
    room against the product's device code on the scene kit's own build of that room."""
    import pytest
    for name, integ in (("whitted", 2), ("direct", 3), ("direct_all", 4)):
        g = np.load(os.path.join(G, f"{name}_96x96_4spp.npz"))
        sk = SceneKit("lights", 96, 96, 4, 31, 2, 0)
        es = emul.scene(sk.desc)
        hits = es.primary_hits(96, 96, 0)
        assert np.mean(hits == g["primary_hit"]) >= 0.999
        img, _ = es.render(RenderParams.make(96, 96, 4, max_depth=int(g["max_depth"]), integrator=integ))
        assert rel_mse(img, g["image"]) <= 1e-3
        es.close(); sk.close()


def test_gaussian_film_golden(emul):
    """Gaussian film: the committed Film::AddSample splat of the reference's samples against the product's gather on the
    scene kit's Cornell box, and GaussianFilter::Evaluate's committed values against the film code's weights."""
    g = np.load(os.path.join(G, "cornell_gaussian_96x96_4spp.npz"))
    sk = SceneKit("cornell", 96, 96, 4, 0, 2, 0)
    es = emul.scene(sk.desc)
    from gnxraytracer_b200.api import FILM_GAUSSIAN, FILM_GAUSSIAN_SUMS
    p = RenderParams.make(96, 96, 4, max_depth=int(g["max_depth"]), film=FILM_GAUSSIAN, filter_radius=float(g["radius"]), filter_alpha=float(g["alpha"]))
    img, _ = es.render(p)
    assert rel_mse(img, g["image"]) <= 1e-3
    p.film = FILM_GAUSSIAN_SUMS
    sums, _ = es.render(p)
    assert np.allclose(sums[..., 3], g["sums"][..., 3], rtol=1e-4), "filter weight sums (scene-kit camera: ulp-level film positions)"
    t = np.load(os.path.join(G, "gaussian_filter_values.npz"))["table"]
    for radius, alpha in {(float(r), float(a)) for r, a in t[:, :2]}:
        rows = t[(t[:, 0] == radius) & (t[:, 1] == alpha)]
        x, y = np.ascontiguousarray(rows[:, 2]), np.ascontiguousarray(rows[:, 3])
        out = np.zeros(x.size, np.float32)
        emul.lib.gnxe_gaussian_eval(radius, alpha, x.size, x.ctypes.data, y.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, rows[:, 4])
    es.close(); sk.close()
