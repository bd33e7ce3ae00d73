#!/bin/bash
# Staged first vertex of the WhittedIntegrator against the per-lane recursion (GNX_WHITTED_STAGED=0), one B200.
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "whitted or Whitted or textures or delta" > $O/r2w_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2w_pytest.log
tail -15 $O/r2w_pytest.log
for w in u1w w1; do
  for s in 1 0; do
    GNX_WHITTED_STAGED=$s python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2w_${w}_s$s.json 2>> $O/r2w.err
    python - $O/r2w_${w}_s$s.json "$w staged=$s" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[2], "failed", e)
PY
  done
done
tail -5 $O/r2w.err
