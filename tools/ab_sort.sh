#!/bin/bash
# Shade queues in slot order (default) against completion order (GNX_SORT_QUEUES=0), one B200.
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
for w in c4 u1w w1; do
  for s in 1 0; do
    GNX_SORT_QUEUES=$s python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2s_${w}_s$s.json 2>> $O/r2s.err
    python - $O/r2s_${w}_s$s.json "$w sorted=$s" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[2], "failed", e)
PY
  done
done
python -m pytest tests -m gpu -x -q -k "parity or fullsize" > $O/r2s_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2s_pytest.log
tail -6 $O/r2s_pytest.log
tail -5 $O/r2s.err
