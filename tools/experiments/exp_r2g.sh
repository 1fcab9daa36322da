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
$T --config c4 --reps 5 > $O/kt_c4.log 2>&1
grep -E "^ +(runs_scan|runs.lookback|runs.lastCTA|union_slices|union_z1|end) " $O/kt_c4.log
python bench.py --no-cpu-baseline --skip-c4 --c3-scans 16 > $O/bench.json 2> $O/bench.err
python - $O/bench.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1])); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2), d['stages_ms'], d['kernels_per_scan'], {k:v.get('ms_per_scan',v.get('ms_per_batch')) for k,v in d['configs'].items() if k!='C5'})
PY
