#!/bin/bash
set -u
O=gpurun_out/r2x
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "c4 or c3 or c2" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for g in 592 1184; do
  echo "== early grid $g"
  MAMRI_STATS_EARLY_CTAS=$g timeout 120 python tools/ktrace.py --config c4 --reps 8 2>&1 | grep -E "^ +(stats|label\.[A-Za-z0-9]+|materialise|end) "
done
for c in c3 c4; do timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
timeout 120 python tools/serial_latency.py --config c4 --conn 26 --reps 30 2>&1 | sed 's/.*bare C ABI/  c4-26 bare/'
