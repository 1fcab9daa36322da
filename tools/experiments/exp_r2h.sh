#!/bin/bash
set -u
O=gpurun_out/r2h
mkdir -p $O
Q="--no-cpu-baseline --skip-c4 --c3-scans 8 --steps 60"
run() { name=$1; shift; env "$@" python bench.py $Q > $O/bench_$name.json 2>> $O/bench.err; python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['ms_per_step'],4), 'lone C2', round(d['configs']['C2']['ms_per_scan'],4), 'C1', round(d['configs']['C1']['ms_per_scan'],4), 'e2e', round(d['e2e']['value'],2))
except Exception as e: print('$name', 'ERR', e)
PY
}
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
run base A=1
run depth3 MAMRI_BENCH_DEPTH=3
run depth4 MAMRI_BENCH_DEPTH=4
run matchains MAMRI_WAVE_MAT_CHAINS=1
run matchains_depth3 MAMRI_WAVE_MAT_CHAINS=1 MAMRI_BENCH_DEPTH=3
run matchains_chains1 MAMRI_WAVE_MAT_CHAINS=1 MAMRI_HBM_CHAINS=1
run ctx4_s4_depth4 MAMRI_BENCH_CONTEXTS=4 MAMRI_BENCH_DEPTH=4
python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4.log 2>&1
grep -E "^ +(runs_scan|runs.lookback|runs.lastCTA|union_slices|end) " $O/kt_c4.log
python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2.log 2>&1
grep -E "^ +(runs_scan|runs.lookback|runs.lastCTA|union_slices|end) " $O/kt_c2.log
for c in c1 c2 c4; do python tools/serial_latency.py --config $c --reps 30; done
