#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
for w in d1 da1; do
  for v in base r2; do
    L=""; [ $v != base ] && L=$PWD/gpurun_in/libgnxrt_$v.so
    GNX_LIB=$L python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2q_${w}_$v.json 2>> $O/r2q.err
    python - $O/r2q_${w}_$v.json "$w $v" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1))
except Exception as e: print(sys.argv[2], "failed", e)
PY
  done
done
tail -3 $O/r2q.err
