#!/bin/bash
set -u
O=gpurun_out/r2k
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for n in 148 296; do
  echo "== SCAN_CTAS=$n"
  for c in c2 c4; do MAMRI_SCAN_CTAS=$n timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
  MAMRI_SCAN_CTAS=$n timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4_scan$n.log 2>&1
  grep -E "^ +(runs_scan|runs.lookback|runs.lastCTA|union_slices|materialise|end) " $O/kt_c4_scan$n.log
done
Q="--no-cpu-baseline --skip-c4 --c3-scans 8 --steps 60"
for mc in 0 1 2 4; do
  MAMRI_WAVE_MID_CHAINS=$mc timeout 300 python bench.py $Q > $O/bench_mid$mc.json 2>> $O/bench.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_mid$mc.json').read().strip().splitlines()[-1]); print('mid_chains=$mc', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2))
except Exception as e: print('mid_chains=$mc', 'ERR', e)
PY
done
timeout 120 python tools/profile_one.py --config c4 --scans 2 > $O/plain_c4.log 2>&1; echo "plain c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_runs_scan|k_stats|k_materialise' -s 3 -c 3 -o $O/full_c4 -f \
    python tools/profile_one.py --config c4 --scans 2 > $O/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
