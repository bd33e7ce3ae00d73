#!/bin/bash
# A/B of the any-hit path on one B200: compressed 8-wide tree (default) against the two-child tree (GNX_ANYHIT_BVH8=0),
# and the second-stream overlap (GNX_ANYHIT_OVERLAP).  Bench lines only (no ncu).
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
python -m pytest tests -m gpu -x -q > $O/r2i_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2i_pytest.log
tail -3 $O/r2i_pytest.log
for w in c2 c1 u1p c3; do
  for v in "1 1" "1 0" "0 0"; do
    set -- $v
    GNX_ANYHIT_BVH8=$1 GNX_ANYHIT_OVERLAP=$2 python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2i_${w}_a$1o$2.json 2>> $O/r2i.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/r2i_${w}_a$1o$2.json").read().strip().splitlines()[-1])
    print("$w any8=$1 overlap=$2", d["ms_per_step"], d["value"], d.get("stages"))
except Exception as e: print("$w $1 $2 failed", e)
PY
  done
done
tail -5 $O/r2i.err
