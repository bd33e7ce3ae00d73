"""Run under torchrun with 2 ranks (tests/test_gpu_multi.py): every rank attaches its single-GPU context to one job
(gnx_comm_attach, the id carried by torch.distributed), renders collectively and rank 0 compares with a plain render."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnxraytracer_b200.api import Context, RenderParams, SceneKit  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h, spp = 200, 120, 9
sk = SceneKit("cornell", w, h, spp, 0, 2, 0)
ctx = Context(local)
box = [ctx.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx.comm_attach(world, rank, box[0])
ctx.upload(sk.desc)
out = np.zeros((h, w, 4), np.float32)
ok = True
for partition in (0, 1):
    p = RenderParams.make(w, h, spp, partition=partition)
    ctx.lib.gnx_render(ctx.h, p, out.ctypes.data if rank == 0 else None, None)
    if rank == 0:
        one = Context(local)
        one.upload(sk.desc)
        ref, _ = one.render(RenderParams.make(w, h, spp))
        one.close()
        good = np.array_equal(out, ref) if partition == 1 else (np.allclose(out[..., :3], ref[..., :3], rtol=2e-6, atol=1e-7) and np.all(out[..., 3] == 1))
        ok = ok and bool(good)
dist.barrier()
if rank == 0:
    print("ATTACH OK" if ok else "ATTACH MISMATCH")
dist.destroy_process_group()
ctx.close()
sys.exit(0 if ok else 1)
