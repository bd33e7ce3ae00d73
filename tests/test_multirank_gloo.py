"""The N > 1 path on CPU: two gloo ranks shard the samples of every pixel, render their range (with
the CPU emulation standing in for the GPU) and sum-reduce the framebuffer; rank 0 must hold the same
image as a single-rank render.  Exercises first_sample / spp_normalize and gnxraytracer_b200.dist."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import _harness
    from gnxraytracer_b200.api import RenderParams, SceneKit
    from gnxraytracer_b200.dist import reduce_framebuffer, sample_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res, spp = 40, 6
    sk = SceneKit("cornell", res, res, spp, 0, 1, 0)
    es = _harness.Emul().scene(sk.desc)
    first, count = sample_range(spp, rank, world)
    img, _ = es.render(RenderParams.make(res, res, count, first_sample=first, spp_normalize=spp))
    fb = torch.from_numpy(img)
    reduce_framebuffer(fb, dst=0)
    # the Gaussian film: ranks exchange the unresolved (sum L f, sum f) buffers
    from gnxraytracer_b200.api import FILM_GAUSSIAN, FILM_GAUSSIAN_SUMS
    from gnxraytracer_b200.dist import reduce_filtered_sums
    sums, _ = es.render(RenderParams.make(res, res, count, first_sample=first, film=FILM_GAUSSIAN_SUMS, filter_radius=2.0, filter_alpha=2.0))
    gs = torch.from_numpy(sums)
    reduce_filtered_sums(gs, dst=0)
    # the library's own job partition (gnx_create_multi / gnx_comm_attach): identical params on every rank, the share is
    # derived from (world, rank) — sample ranges, and interleaved 32 x 32 tiles on a ragged 70 x 45 image
    job = RenderParams.make(res, res, spp, partition=0)
    share = torch.from_numpy(es.render_share(job, world, rank))
    dist.reduce(share, dst=0, op=dist.ReduceOp.SUM)
    sk2 = SceneKit("cornell", 70, 45, 3, 0, 1, 0)
    es2 = _harness.Emul().scene(sk2.desc)
    tjob = RenderParams.make(70, 45, 3, partition=1)
    tshare = torch.from_numpy(es2.render_share(tjob, world, rank))
    dist.reduce(tshare, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        full, _ = es.render(RenderParams.make(res, res, spp))
        gfull, _ = es.render(RenderParams.make(res, res, spp, film=FILM_GAUSSIAN, filter_radius=2.0, filter_alpha=2.0))
        tfull, _ = es2.render(RenderParams.make(70, 45, 3))
        np.savez(out_path, reduced=fb.numpy(), full=full, greduced=gs.numpy(), gfull=gfull, share=share.numpy(), tshare=tshare.numpy(), tfull=tfull)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reduce_to_the_single_rank_image(tmp_path, emul):
    out = str(tmp_path / "out.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    d = np.load(out)
    assert np.allclose(d["reduced"][..., :3], d["full"][..., :3], rtol=1e-5, atol=1e-6)
    assert np.all(d["reduced"][..., 3] == 1.0)
    assert np.allclose(d["greduced"][..., :3], d["gfull"][..., :3], rtol=1e-4, atol=1e-6)
    assert np.all(d["greduced"][..., 3] == 1.0)
    # the library's partitions: sample ranges to float rounding, tiles bit for bit
    assert np.allclose(d["share"][..., :3], d["full"][..., :3], rtol=1e-5, atol=1e-6) and np.all(d["share"][..., 3] == 1.0)
    assert np.array_equal(d["tshare"], d["tfull"])


def test_sample_ranges_partition():
    from gnxraytracer_b200.dist import sample_range, weak_sample_range
    for spp in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            r = [sample_range(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and sum(c for _, c in r) == spp
            for (f0, c0), (f1, _) in zip(r, r[1:]):
                assert f0 + c0 == f1
    assert weak_sample_range(64, 3) == (192, 64)


def test_tile_partition_covers_every_pixel_once(emul):
    """gnx_path.cuh pixel_xy / local_tile_count: for 1..8 devices and ragged image sizes the devices' tiles partition the image."""
    from gnxraytracer_b200.api import RenderParams, SceneKit
    import _harness
    sk = SceneKit("cornell", 8, 8, 1, 0, -1, 0)
    es = _harness.Emul().scene(sk.desc)
    for (w, h) in ((33, 65), (96, 32), (100, 7)):
        p = RenderParams.make(w, h, 1, max_depth=0, partition=1, light_strategy=0)
        for n in (1, 2, 3, 5, 8):
            cover = np.zeros((h, w), np.int32)
            for g in range(n):
                cover += (es.render_share(p, n, g)[..., 3] == 1.0).astype(np.int32)
            assert np.all(cover == 1), (w, h, n)
