#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
Q="--no-cpu-baseline --no-bridge --no-strong-record"
python -m pytest tests -m gpu -x -q -k "volpath or smoke or fog or textures or fullsize or device_built" > $O/r2v_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2v_pytest.log
tail -5 $O/r2v_pytest.log
python bench.py --workload c4 --steps 5 --warmup 3 $Q > $O/r2v_c4.json 2>> $O/r2v.err
python - $O/r2v_c4.json "c4 record layout" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d["ms_per_step"],3), round(d["value"],1), [(s["kernel"], round(s["ms"],1), round(s["frac"],3)) for s in d["roofline"]["stages"]])
PY
tail -3 $O/r2v.err
