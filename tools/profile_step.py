#!/usr/bin/env python
"""Runs `--steps` renders of a bench workload with nothing else around them (for ncu / quick timing).
Prints the library's own per-stage CUDA-event times and counters of the last step."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (bench.py points fd 1 at stderr for its one-JSON-line contract: undo that here)
from bench import WORKLOADS  # noqa: E402
os.dup2(bench._REAL_STDOUT, 1)
from gnxraytracer_b200.api import Context, RenderParams, SceneKit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--res", type=int, default=0)
ap.add_argument("--batch-spp", type=int, default=0)
a = ap.parse_args()
wl = WORKLOADS[a.workload]
scene, (p0, p1, p2), W, H, spp, depth, desc = wl["scene"], wl["args"], wl["w"], wl["h"], wl["spp"], wl["depth"], wl["desc"]
if a.spp:
    spp = a.spp
if a.res:
    W = H = a.res
ctx = Context(0)
sk = SceneKit(scene, W, H, spp, p0, p1, p2)
ctx.upload(sk.desc)
p = RenderParams.make(W, H, spp, max_depth=depth, batch_spp=a.batch_spp, integrator=wl["integ"])
for i in range(a.steps):
    img, st = ctx.render(p)
d = st.as_dict()
d["mpaths_per_s"] = st.paths / st.device_ms / 1e3
d["mrays_per_s"] = st.rays / st.device_ms / 1e3
d["nodes_per_ray"] = st.nodes_visited / st.rays
d["tris_per_ray"] = st.tris_tested / st.rays
d["mean"] = float(img[..., :3].mean())
print(json.dumps(d))
