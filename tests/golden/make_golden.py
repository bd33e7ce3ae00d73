"""Generates the committed golden fixtures from the UNMODIFIED reference (needs oracle/_ref, i.e.
/root/reference at build time).  Run from the repo root:  python tests/golden/make_golden.py

  halton_values.npz        HaltonSampler::SampleDimension for a table of (index, dim)      [bit-exact]
  cornell_96x96_4spp.npz   reference Render() image + per-pixel primary-hit primitive ids
  dragon_96x96_4spp.npz    same for the (256 x 32)-quad dragon-class mesh under MonValley
  whitted_96x96_4spp.npz   WhittedIntegrator::Render on the lights room (area + point + spot + distant + skybox)
  direct_96x96_4spp.npz    DirectLightingIntegrator(UniformSampleOne)::Render on the same room
  direct_all_96x96_4spp.npz  DirectLightingIntegrator(UniformSampleAll)::Render on the same room
  cornell_gaussian_96x96_4spp.npz  Film::AddSample splat of the reference's camera samples / Li with GaussianFilter(2, 2)::Evaluate
                           (oracle/ref_harness.cpp::gnxh_reference_gaussian_film): image + (sum L f, sum f)
  gaussian_filter_values.npz  GaussianFilter::Evaluate for a table of (radius, alpha, x, y)                  [<= 2 ulp: expf]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _harness  # noqa: E402


def main():
    ref = _harness.Ref()
    rs = ref.scene("cornell", 96, 96, 4)
    rng = np.random.default_rng(2026)
    idx = rng.integers(0, 31104 * 1024 + 31103, 4096).astype(np.int64)
    dim = rng.integers(0, 1000, 4096).astype(np.int32)
    dim[:512] = rng.integers(0, 6, 512)
    np.savez_compressed(os.path.join(HERE, "halton_values.npz"), index=idx, dim=dim, value=rs.sample_dims(idx, dim),
                        pixel_index=np.array([[x, y, s, rs.sample_index(x, y, s)] for x, y, s in rng.integers(0, 96, (256, 3))], np.int64))
    for preset, name in (("cornell", "cornell_96x96_4spp"), ("dragon", "dragon_96x96_4spp"), ("whitted", "whitted_96x96_4spp"),
                         ("direct", "direct_96x96_4spp"), ("direct_all", "direct_all_96x96_4spp")):
        rs = ref.scene(preset, 96, 96, 4)
        img, _ = rs.render_reference(max_depth=5)
        px, py = _harness.grid(96, 96)
        _, prim = rs.reference_samples(px, py, np.zeros(px.size, np.int32), want_rgb=False)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=img.astype(np.float16).astype(np.float32) if False else img,
                            primary_hit=prim.astype(np.int32), max_depth=5, spp=4)
        print(name, img[..., :3].mean())
    rs = ref.scene("cornell", 96, 96, 4)
    img, sums = rs.reference_gaussian_film(2.0, 2.0, max_depth=5)
    np.savez_compressed(os.path.join(HERE, "cornell_gaussian_96x96_4spp.npz"), image=img, sums=sums, radius=2.0, alpha=2.0, max_depth=5, spp=4)
    tabs = []
    for radius, alpha in ((2.0, 2.0), (1.5, 0.5), (3.0, 1.0)):
        x = rng.uniform(-radius, radius, 1024).astype(np.float32)
        y = rng.uniform(-radius, radius, 1024).astype(np.float32)
        v = np.zeros(1024, np.float32)
        ref.lib.gnxh_reference_gaussian_eval(radius, alpha, 1024, x.ctypes.data, y.ctypes.data, v.ctypes.data)
        tabs.append(np.stack([np.full(1024, radius, np.float32), np.full(1024, alpha, np.float32), x, y, v], 1))
    np.savez_compressed(os.path.join(HERE, "gaussian_filter_values.npz"), table=np.concatenate(tabs))


if __name__ == "__main__":
    main()
