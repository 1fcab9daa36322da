#!/bin/bash
set -u
O=gpurun_out/${1:-r2d}
mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
T="python tools/ktrace.py --reps 20"
$T --config c2 > $O/kt_c2.log 2>&1; echo "kt c2 rc=$?"
$T --config c1 > $O/kt_c1.log 2>&1
$T --config c3 > $O/kt_c3.log 2>&1
$T --config c4 --reps 5 > $O/kt_c4.log 2>&1
$T --config c4 --conn 26 --reps 5 > $O/kt_c4_26.log 2>&1
MAMRI_CLOSE_SPECIALISE=0 $T --config c2 > $O/kt_c2_nospec.log 2>&1
for c in c1 c2 c3 c4; do python tools/serial_latency.py --config $c --reps 30 > $O/serial_$c.log 2>&1; done
python bench.py --no-cpu-baseline --skip-c4 --c3-scans 16 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/serial_*.log
python tools/profile_one.py --scans 3 > $O/plain.log 2>&1; echo "plain rc=$?"
if [ "${NCU:-0}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:'k_close_fused|k_runs_scan|k_union|k_flatten|k_select|k_stats' -s 16 -c 8 \
    -o $O/full_mid -f python tools/profile_one.py --scans 3 > $O/ncu_full_mid.log 2>&1; echo "ncu mid rc=$?"
fi
python - $O/bench.json <<'PY'
import json
import sys; d=json.load(open(sys.argv[1])); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2), d['stages_ms'], d['kernels_per_scan'], {k:v.get('ms_per_scan',v.get('ms_per_batch')) for k,v in d['configs'].items() if k!='C5'})
PY
