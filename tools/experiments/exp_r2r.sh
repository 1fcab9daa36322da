#!/bin/bash
# the early sums of a noisy scan (k_stats<1>, now light) beside `materialise` instead of before it; source-level profile of it
set -u
O=gpurun_out/r2r
mkdir -p $O
echo "== default (sums before materialise)"
timeout 120 python tools/serial_latency.py --config c4 --reps 30 2>&1 | sed 's/.*bare C ABI/  c4 bare/'
for g in 148 296 592 1184; do
  echo "== STATS_SPLIT=3 early grid $g"
  MAMRI_STATS_SPLIT=3 MAMRI_STATS_EARLY_CTAS=$g timeout 120 python tools/serial_latency.py --config c4 --reps 30 2>&1 | sed 's/.*bare C ABI/  c4 bare/'
done
MAMRI_STATS_SPLIT=3 MAMRI_STATS_EARLY_CTAS=148 timeout 120 python tools/serial_latency.py --config c4 --conn 26 --reps 30 2>&1 | sed 's/.*bare C ABI/  c4-26 bare (split 3, 148)/'
MAMRI_STATS_SPLIT=3 MAMRI_STATS_EARLY_CTAS=148 timeout 120 python tools/ktrace.py --config c4 --reps 10 > $O/kt_c4_split3.log 2>&1; cat $O/kt_c4_split3.log | grep -E "^ +(select|select.end.last|stats|materialise|stats.finalise|final|end) "
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_stats' -s 2 -c 1 -o $O/full_c4_stats -f \
    python tools/profile_one.py --config c4 --scans 2 > $O/ncu_full_c4.log 2>&1; echo "ncu full c4 stats rc=$?"
