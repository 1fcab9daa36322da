#!/bin/bash
set -u
O=gpurun_out/r2l
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for sp in 1 0 2; do
  echo "== STATS_SPLIT=$sp"
  for c in c1 c2 c3 c4; do MAMRI_STATS_SPLIT=$sp timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
done
timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4.log 2>&1; cat $O/kt_c4.log | grep -E "^ +(threshold.lastCTA|close|erode|runs_scan|runs.lastCTA|union_slices|union_z1|union_z2|flatten_rank|select|stats|materialise|stats.finalise|final|end) "
Q="--no-cpu-baseline --skip-c4 --c3-scans 8 --steps 60"
for mc in 0 4 3 5 6 0 4; do
  MAMRI_WAVE_MID_CHAINS=$mc timeout 300 python bench.py $Q > $O/bench_mid$mc.json 2>> $O/bench.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_mid$mc.json').read().strip().splitlines()[-1]); print('mid_chains=$mc', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2))
except Exception as e: print('mid_chains=$mc', 'ERR', e)
PY
done
MAMRI_WAVE_MID_CHAINS=4 MAMRI_STATS_SPLIT=2 timeout 300 python bench.py $Q > $O/bench_mid4_split2.json 2>> $O/bench.err; tail -c 3000 $O/bench_mid4_split2.json | grep -o '"value": [0-9.]*' | head -1
