#!/bin/bash
# A/B of the closest-hit path on one B200: compressed 8-wide tree + retrace of flagged rays (default) against the
# reference-order two-child tree (GNX_CLOSEST_BVH8=0); resident-block variants of the wide kernels (GNX_LIB).
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
L=$PWD/gnxraytracer_b200/lib
line() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[2], "failed", e)
PY
}
for w in c2 c1 u1p c3; do
  GNX_CLOSEST_BVH8=0 python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2j_${w}_c0.json 2>> $O/r2j.err; line $O/r2j_${w}_c0.json "$w closest8=0"
  python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2j_${w}_c1.json 2>> $O/r2j.err; line $O/r2j_${w}_c1.json "$w closest8=1 b8"
  for b in 6; do
    GNX_LIB=$L/libgnxrt_b$b.so python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2j_${w}_c1b$b.json 2>> $O/r2j.err; line $O/r2j_${w}_c1b$b.json "$w closest8=1 b$b"
  done
done
python -m pytest tests -m gpu -x -q --durations=12 > $O/r2j_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2j_pytest.log
tail -25 $O/r2j_pytest.log
tail -5 $O/r2j.err
