"""HaltonSampler parity (bit-exact): the device sampler functions, compiled for the host, against the
reference's own HaltonSampler::SampleDimension / GetIndexForSample (samplers/HaltonSampler.cpp:63-94)."""
import numpy as np
import pytest

from gnxraytracer_b200.api import SceneKit


@pytest.mark.parametrize("res", [(64, 64), (200, 120), (37, 513)])
def test_sample_dimensions_bit_exact(ref, emul, res):
    w, h = res
    rs = ref.scene("cornell", w, h, 16)
    es = emul.scene(rs.desc)
    rng = np.random.default_rng(7)
    n = 50000
    idx = rng.integers(0, 31104 * 1024 + 31103, n).astype(np.int64)
    idx[:64] = np.arange(64)
    dim = rng.integers(0, 1000, n).astype(np.int32)
    dim[:2000] = rng.integers(0, 8, 2000)
    a, b = rs.sample_dims(idx, dim), es.sample_dims(idx, dim)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert a.min() >= 0 and a.max() < 1
    for x, y, s in rng.integers(0, min(w, h), (300, 3)):
        assert rs.sample_index(x, y, s) == es.sample_index(x, y, s)
    rs.close(); es.close()


def test_library_generated_permutations_match_reference(ref, emul):
    """A scene-kit scene carries no permutation table: the library derives it from a default-seeded
    PCG32 (samplers/HaltonSampler.cpp:36-39).  The values must equal the reference's table."""
    rs = ref.scene("cornell", 64, 64, 4)
    sk = SceneKit("cornell", 64, 64, 4, 0, 2, 0)
    es = emul.scene(sk.desc)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 1 << 25, 40000).astype(np.int64)
    dim = rng.integers(2, 1000, 40000).astype(np.int32)
    a, b = rs.sample_dims(idx, dim), es.sample_dims(idx, dim)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    rs.close(); es.close(); sk.close()
