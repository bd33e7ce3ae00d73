#!/bin/bash
# Round-2 closing run on one B200: GPU tests, one bench line per workload (with the CPU reference and the bridge timing),
# the reference arm, then the ncu evidence (tools/final_evidence.sh).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=8 > $O/r2z_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2z_pytest.log
tail -4 $O/r2z_pytest.log
python bench.py --steps 10 --warmup 3 > $O/r2z_bench_c2.json 2> $O/r2z_bench_c2.err
for w in c1 c3 c4 u1p u1w w1 d1 da1; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-strong-record > $O/r2z_bench_$w.json 2> $O/r2z_bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 0 > $O/r2z_bench_c2_reference.json 2> $O/r2z_bench_ref.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-bridge --no-strong-record --film gaussian > $O/r2z_bench_c2_gaussian.json 2>> $O/r2z_bench_c2.err
for f in $O/r2z_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("bench_")[1], d.get("ms_per_step"), d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
bash tools/final_evidence.sh
