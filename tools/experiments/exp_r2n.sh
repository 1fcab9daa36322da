#!/bin/bash
set -u
O=gpurun_out/r2n
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for sp in 1 0; do
  echo "== STATS_SPLIT=$sp"
  for c in c1 c2 c3 c4; do MAMRI_STATS_SPLIT=$sp timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
done
for c in c4 c2 c1; do
timeout 120 python tools/ktrace.py --config $c --reps 10 > $O/kt_$c.log 2>&1; echo "== $c"; cat $O/kt_$c.log | grep -E "^ +(threshold.lastCTA|close|erode|runs_scan|runs.lastCTA|union_slices|union_z1|union_z2|flatten_rank|select|stats|materialise|stats.finalise|final|end) "
done
MAMRI_STATS_SPLIT=0 timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4_nosplit.log 2>&1; echo "== c4 nosplit"; cat $O/kt_c4_nosplit.log | grep -E "^ +(select|stats|materialise|stats.finalise|final|end) "
Q="--no-cpu-baseline --skip-c4 --c3-scans 8 --steps 60"
timeout 300 python bench.py $Q > $O/bench.json 2>> $O/bench.err
python - <<PY
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print('batch', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2), 'C2', d['configs']['C2']['ms_per_scan'], 'C1', d['configs']['C1']['ms_per_scan'])
PY
