#!/bin/bash
set -u
O=gpurun_out/r2p
mkdir -p $O
for g in 148 296 592 1184; do
  MAMRI_STATS_EARLY_CTAS=$g timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4_g$g.log 2>&1; echo "== c4 early grid $g: $(head -1 $O/kt_c4_g$g.log | sed 's/.*median/median/')"; cat $O/kt_c4_g$g.log | grep -E "^ +(stats|materialise|end) "
done
timeout 120 python tools/profile_one.py --config c4 --scans 2 > $O/plain_c4.log 2>&1; echo "plain c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stats' -s 2 -c 2 -o $O/full_c4_stats -f \
    python tools/profile_one.py --config c4 --scans 2 > $O/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
