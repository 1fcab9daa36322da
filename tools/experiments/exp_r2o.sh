#!/bin/bash
set -u
O=gpurun_out/r2o
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for sp in 1 0 2; do
  echo "== STATS_SPLIT=$sp"
  for c in c2 c3 c4; do MAMRI_STATS_SPLIT=$sp timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
done
timeout 120 python tools/ktrace.py --config c4 --reps 10 > $O/kt_c4.log 2>&1; echo "== c4"; cat $O/kt_c4.log | grep -E "^ +(select|select.end.last|stats|materialise|stats.finalise|final|end) "
timeout 120 python tools/ktrace.py --config c4 --conn 26 --reps 10 > $O/kt_c4_26.log 2>&1; echo "== c4/26"; head -1 $O/kt_c4_26.log
