"""N-device jobs behind the C ABI and the drop-in class (run on the B200 box with `-m gpu`).

gnx_create_multi drives several GPUs from ONE process — what a reference app calling integrator->Render(scene, t)
(ui/RenderThread.cpp:169-175) needs to use a whole box.  On a one-GPU box the same code runs with device 0 listed more
than once (the shares time-share the GPU, the reduce is the one-kernel peer sum); with >= 2 GPUs visible the real
devices are used as well (NCCL or peer-to-peer loads over NVLink).  The multi-PROCESS variant (gnx_comm_attach, what
bench.py does under torchrun) needs two GPUs and is skipped otherwise.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from _harness import rel_mse
from gnxraytracer_b200.api import (FILM_GAUSSIAN, PARTITION_SAMPLES, PARTITION_TILES, Context, GnxError, RenderParams, SceneKit,
                                   load_library)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_sets():
    n = load_library().gnx_device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1], list(range(min(n, 8)))]
    return sets


@pytest.mark.parametrize("name,args,integ,depth", [("cornell", (0, 2, 0), 0, 5), ("lights", (31, 2, 2), 2, 4), ("smoke", (0, 0, 0), 1, 4)])
def test_tile_partition_is_bit_equal_and_sample_partition_equal_to_rounding(name, args, integ, depth):
    w, h, spp = 200, 120, 8   # ragged: edge tiles in both directions
    sk = SceneKit(name, w, h, spp, *args)
    one = Context(0)
    one.upload(sk.desc)
    base = dict(max_depth=depth, integrator=integ)
    ref_img, ref_st = one.render(RenderParams.make(w, h, spp, **base))
    for devices in _device_sets():
        m = Context(devices=devices)
        assert m.num_devices == len(devices)
        m.upload(sk.desc)
        tiles, st = m.render(RenderParams.make(w, h, spp, partition=PARTITION_TILES, **base))
        assert np.array_equal(tiles, ref_img), f"{devices}: the tile partition must reproduce the single-device image bit for bit"
        assert st.paths == ref_st.paths and st.rays == ref_st.rays
        smp, st2 = m.render(RenderParams.make(w, h, spp, partition=PARTITION_SAMPLES, **base))
        assert np.allclose(smp[..., :3], ref_img[..., :3], rtol=2e-6, atol=1e-7), f"{devices}: sample ranges regroup the float sum only"
        assert np.all(smp[..., 3] == 1.0) and st2.paths == ref_st.paths and st2.rays == ref_st.rays
        again, _ = m.render(RenderParams.make(w, h, spp, partition=PARTITION_SAMPLES, **base))
        assert np.array_equal(again, smp), "deterministic"
        m.close()
    one.close(); sk.close()


def test_fewer_samples_than_devices_and_gaussian_film():
    w, h = 96, 64
    sk = SceneKit("cornell", w, h, 8, 0, 2, 0)
    one = Context(0)
    one.upload(sk.desc)
    m = Context(devices=[0, 0, 0])
    m.upload(sk.desc)
    a, _ = one.render(RenderParams.make(w, h, 2))
    b, _ = m.render(RenderParams.make(w, h, 2))          # 2 samples over 3 shares: one share is empty
    assert np.allclose(a, b, rtol=2e-6, atol=1e-7)
    g1, _ = one.render(RenderParams.make(w, h, 8, film=FILM_GAUSSIAN, filter_radius=2.0, filter_alpha=2.0))
    g3, _ = m.render(RenderParams.make(w, h, 8, film=FILM_GAUSSIAN, filter_radius=2.0, filter_alpha=2.0))   # sums reduced, resolved on the root
    assert np.allclose(g1, g3, rtol=1e-5, atol=1e-6)
    with pytest.raises(GnxError) as e:                     # a filter needs neighbours across tiles
        m.render(RenderParams.make(w, h, 8, film=FILM_GAUSSIAN, filter_radius=2.0, filter_alpha=2.0, partition=PARTITION_TILES))
    assert e.value.code == -4
    one.close(); m.close(); sk.close()


def test_framebuffer_sink_running_mean_and_8bit_image():
    """gnx_render_framebuffer: FrameBuffer::update_f_u_c on the device (ui/FrameBuffer.h:127-149)."""
    w, h, spp = 96, 80, 4
    sk = SceneKit("cornell", w, h, spp, 0, 2, 0)
    ctx = Context(0)
    ctx.upload(sk.desc)
    img, _ = ctx.render(RenderParams.make(w, h, spp))
    f = np.full((h, w, 4), 7.0, np.float32)   # stale content: the first pass forgets the colours, keeps the alpha
    u = np.zeros((h, w, 4), np.uint8)
    ctx.render_framebuffer(RenderParams.make(w, h, spp), 1, f, u)
    assert np.array_equal(f[..., :3], img[..., :3]) and np.all(f[..., 3] == 7.0)
    want = ((1.0 - np.exp(-f[..., :3] / np.float32(0.25))) * 255).astype(np.uint8)
    assert np.abs(u[..., :3].astype(int) - want.astype(int)).max() <= 1 and np.all(u[..., 3] == 255)
    # second pass with other samples: mean of the two passes
    other, _ = ctx.render(RenderParams.make(w, h, spp, first_sample=spp))
    ctx.render_framebuffer(RenderParams.make(w, h, spp, first_sample=spp), 2, f, u)
    assert np.allclose(f[..., :3], 0.5 * other[..., :3] + 0.5 * img[..., :3], rtol=1e-6, atol=1e-7)
    # a context that has not seen the earlier passes takes the running mean from the caller's buffer
    fresh = Context(0)
    fresh.upload(sk.desc)
    f2 = np.zeros((h, w, 4), np.float32)
    f2[..., :3] = img[..., :3]
    fresh.render_framebuffer(RenderParams.make(w, h, spp, first_sample=spp), 2, f2, u)
    assert np.allclose(f2[..., :3], f[..., :3], rtol=1e-6, atol=1e-7)
    ctx.close(); fresh.close(); sk.close()


def test_bridge_render_fills_the_framebuffer_like_the_reference(ref):
    """Integrator::Render is the boundary: three passes of the UI's loop through the drop-in class against three passes of
    the reference's own Render — float buffer, 8-bit buffer (+-1: expf), both alphas."""
    w, h, spp = 100, 72, 4
    rs = ref.scene("cornell", w, h, spp)
    f_ref, u_ref = rs.render_reference_passes(3)
    f, u, wall = rs.render_cuda_passes(3)
    assert rel_mse(f, f_ref) <= 1e-6 and wall > 0
    assert np.array_equal(f[..., 3], f_ref[..., 3]), "the float alpha is never written by Render (core/Integrator.cpp:307-310)"
    d = np.abs(u.astype(int) - u_ref.astype(int))
    assert d.max() <= 1 and np.mean(d > 0) < 0.02, f"8-bit image: max diff {d.max()}, {np.mean(d > 0)} of the bytes differ"
    assert np.all(u[..., 3] == 255)
    # progressive passes: pass k renders samples [k spp, (k+1) spp) — the running mean equals ONE render of 3 spp samples
    rs.set_progressive(True)
    fp, up, _ = rs.render_cuda_passes(3)
    rs.set_progressive(False)
    ctx = Context(0)
    ctx.upload(rs.desc)
    full, _ = ctx.render(RenderParams.make(w, h, 3 * spp))
    assert np.allclose(fp[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)
    ctx.close(); rs.close()


def test_bridge_on_several_devices(ref):
    """CUDAPathIntegrator(..., devices): one Render() call on N GPUs; tiles bit-equal to one GPU, sample ranges to rounding."""
    w, h, spp = 160, 96, 8
    rs = ref.scene("dragon", w, h, spp)
    img1, _, st1 = rs.render_cuda(max_depth=5)
    img_ref, _ = rs.render_reference(max_depth=5)
    for devices in _device_sets():
        rs.set_devices(devices, PARTITION_TILES)
        a, _, st = rs.render_cuda(max_depth=5)
        assert np.array_equal(a, img1) and st.paths == st1.paths, devices
        rs.set_devices(devices, PARTITION_SAMPLES)
        b, _, _ = rs.render_cuda(max_depth=5)
        assert np.allclose(b, img1, rtol=2e-6, atol=1e-7) and rel_mse(b, img_ref) <= 1e-3, devices
    rs.set_devices([], 0)
    rs.close()


def test_attached_ranks_reduce_inside_the_library():
    """gnx_comm_attach: two processes, one GPU each, identical params on both; rank 0 receives the image."""
    if load_library().gnx_device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + (os.getpid() % 200)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "mp_attach_check.py")], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "ATTACH OK" in r.stdout
