#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
for w in c4; do
  for s in 0 1; do
    GNX_L2_ENV=$s python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2l_${w}_e$s.json 2>> $O/r2l.err
    python - $O/r2l_${w}_e$s.json "$w l2_env=$s" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[2], "failed", e)
PY
  done
done

tail -3 $O/r2l.err
