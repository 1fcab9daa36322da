#!/bin/bash
set -u
O=gpurun_out/r2g
mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
T="python tools/ktrace.py --reps 20"
$T --config c2 > $O/kt_c2.log 2>&1
MAMRI_CLOSE_SPECIALISE=1 $T --config c2 > $O/kt_c2_spec.log 2>&1
grep -E "^ +(close|close.dilated|close.eroded|close.lastCTA|runs_scan|runs.lastCTA|union_slices|end) " $O/kt_c2.log $O/kt_c2_spec.log
for c in c1 c2 c4; do python tools/serial_latency.py --config $c --reps 30 > $O/serial_$c.log 2>&1; done
cat $O/serial_*.log
