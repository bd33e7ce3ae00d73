import sys
sys.path.insert(0, ".")
from gnxraytracer_b200.api import Context, RenderParams, SceneKit
ctx = Context(0)
sk = SceneKit("dragon", 1024, 1024, 64, 0, 0, 0)
p = RenderParams.make(1024, 1024, 64, max_depth=5)
sk.strip_bvh()
ctx.upload(sk.desc)
for i in range(3): img2, st2 = ctx.render(p)
print("device LBVH: build %.2f ms, render %.2f ms, nodes/ray %.1f tris/ray %.2f" % (ctx.bvh_build_ms, st2.device_ms, st2.nodes_visited / st2.rays, st2.tris_tested / st2.rays))
