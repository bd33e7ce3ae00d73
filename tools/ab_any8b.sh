#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
for w in c2 u1p u1w c3; do
  for b in 8 6 5; do
    L=""; [ $b != 8 ] && L=$PWD/gpurun_in/libgnxrt_a$b.so
    GNX_LIB=$L python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2b_${w}_a$b.json 2>> $O/r2b.err
    python - $O/r2b_${w}_a$b.json "$w any8 blocks=$b" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[2], "failed", e)
PY
  done
done
tail -3 $O/r2b.err
