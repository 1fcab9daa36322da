#!/bin/bash
set -u
O=gpurun_out/r2c
mkdir -p $O
./tools/micro/latency > $O/latency.log 2>&1; echo "latency rc=$?"
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
T="python tools/ktrace.py --reps 20"
$T --config c2 > $O/kt_c2.log 2>&1; echo "kt c2 rc=$?"
$T --config c1 > $O/kt_c1.log 2>&1
$T --config c3 > $O/kt_c3.log 2>&1
MAMRI_CLOSE_CTAS_PER_SM=1 $T --config c2 > $O/kt_c2_close1.log 2>&1
MAMRI_CLOSE_CTAS_PER_SM=2 MAMRI_CLOSE_ZSD=2 MAMRI_CLOSE_ZSE=2 $T --config c2 > $O/kt_c2_close_zs2.log 2>&1
MAMRI_CLOSE_TZ=8 $T --config c2 > $O/kt_c2_close_tz8.log 2>&1
MAMRI_CLOSE_TY=8 MAMRI_CLOSE_TZ=8 $T --config c2 > $O/kt_c2_close_8_8.log 2>&1
MAMRI_STREAM_HINTS=0 $T --config c2 > $O/kt_c2_nohints.log 2>&1
$T --config c4 --reps 5 > $O/kt_c4.log 2>&1
for c in c1 c2 c3 c4; do python tools/serial_latency.py --config $c --reps 30 > $O/serial_$c.log 2>&1; done
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
Q="--no-cpu-baseline --skip-c4 --c3-scans 16"
MAMRI_CLOSE_FUSED=0 python bench.py $Q > $O/bench_unfused.json 2>> $O/bench.err
MAMRI_THR_V8=0 python bench.py $Q > $O/bench_v4.json 2>> $O/bench.err
MAMRI_CLOSE_FUSED=0 MAMRI_THR_V8=0 python bench.py $Q > $O/bench_unfused_v4.json 2>> $O/bench.err
MAMRI_CLOSE_CTAS_PER_SM=1 python bench.py $Q > $O/bench_close1.json 2>> $O/bench.err
MAMRI_BENCH_BODY_U8=1 python bench.py $Q > $O/bench_u8.json 2>> $O/bench.err
tail -5 $O/bench.err
cat $O/serial_*.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/bench*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2), 'ceil', round(d['e2e']['copy_ceiling']['value'],2), d['stages_ms'], d['kernels_per_scan'], {k:(round(v.get('ms_per_scan',v.get('ms_per_batch',0)),4), v.get('parity',{}).get('labels_bit_exact'), v.get('gathered_equals_single_gpu')) for k,v in d['configs'].items() if k!='C5'}, d['configs'].get('C5'))
    except Exception as e: print(f, 'ERR', e)
PY
