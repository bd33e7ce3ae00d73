#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
for v in base m2v m2; do
  L=""; [ $v != base ] && L=$PWD/gpurun_in/libgnxrt_$v.so
  GNX_LIB=$L python bench.py --workload c4 --steps 5 --warmup 3 $Q > $O/r2p_c4_$v.json 2>> $O/r2p.err
  python - $O/r2p_c4_$v.json "c4 $v" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), [(s["kernel"][11:18], round(s["ms"],1)) for s in d["roofline"]["stages"]])
except Exception as e: print(sys.argv[2], "failed", e)
PY
done
tail -3 $O/r2p.err
