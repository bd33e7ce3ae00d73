#!/bin/bash
# Experiment: extension queue re-binned by origin Morton code + direction octant between bounces (GNX_REBIN=1), one B200.
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
for w in u1p c1 c2; do
  for s in 0 1; do
    GNX_REBIN=$s python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2r_${w}_r$s.json 2>> $O/r2r.err
    python - $O/r2r_${w}_r$s.json "$w rebin=$s" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)}, "mean", d.get("image_mean"))
except Exception as e: print(sys.argv[2], "failed", e)
PY
  done
done
python - <<'PY'
import os, sys
sys.path.insert(0, "tests")
import numpy as np
from gnxraytracer_b200.api import Context, RenderParams, SceneKit
sk = SceneKit("ui", 160, 160, 4, 0, 0, 0)
p = RenderParams.make(160, 160, 4, max_depth=15)
out = []
for r in ("0", "1"):
    os.environ["GNX_REBIN"] = r
    c = Context(0); c.upload(sk.desc); img, st = c.render(p); out.append(img); c.close()
print("rebin bit-equal:", np.array_equal(out[0], out[1]))
PY
tail -3 $O/r2r.err
