#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
python -m pytest tests -m gpu -x -q > $O/r2y_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2y_pytest.log
tail -5 $O/r2y_pytest.log
for w in c2 c1 c3 c4 u1p u1w w1; do
  python bench.py --workload $w --steps 5 --warmup 3 $Q > $O/r2y_$w.json 2>> $O/r2y.err
  python - $O/r2y_$w.json "$w" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(sys.argv[2], "failed", e)
PY
done
python bench.py --steps 5 --warmup 3 $Q --film gaussian > $O/r2y_c2_gauss.json 2>> $O/r2y.err
python - $O/r2y_c2_gauss.json "c2 gaussian" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)})
PY
tail -3 $O/r2y.err
