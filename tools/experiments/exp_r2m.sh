#!/bin/bash
set -u
O=gpurun_out/r2m
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for sp in 1 0 2; do
  echo "== STATS_SPLIT=$sp"
  for c in c1 c2 c3 c4; do MAMRI_STATS_SPLIT=$sp timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
done
echo "== c4 conn 26"; timeout 120 python tools/serial_latency.py --config c4 --conn 26 --reps 20 2>&1 | sed 's/.*bare C ABI/  c4 bare/'
timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4.log 2>&1; cat $O/kt_c4.log | grep -E "^ +(threshold.lastCTA|close|erode|runs_scan|runs.lastCTA|union_slices|union_z1|union_z2|flatten_rank|select|stats|materialise|stats.finalise|final|end) "
timeout 120 python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2.log 2>&1; cat $O/kt_c2.log | grep -E "^ +(select|stats|materialise|stats.finalise|final|end) "
MAMRI_STATS_SPLIT=2 timeout 120 python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2_split2.log 2>&1; cat $O/kt_c2_split2.log | grep -E "^ +(select|stats|materialise|stats.finalise|final|end) "
