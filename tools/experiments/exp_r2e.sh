#!/bin/bash
# batch-throughput sweep: which knob moves the 8-scan wave
set -u
O=gpurun_out/r2e
mkdir -p $O
Q="--no-cpu-baseline --skip-c4 --c3-scans 8 --steps 60"
run() { name=$1; shift; env "$@" python bench.py $Q > $O/bench_$name.json 2>> $O/bench.err; python - <<PY
import json
try:
    d=json.load(open('$O/bench_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],4), 'lone C2', round(d['configs']['C2']['ms_per_scan'],4), 'e2e', round(d['e2e']['value'],2))
except Exception as e: print('$name', 'ERR', e)
PY
}
run base A=1
run ctx4 MAMRI_BENCH_CONTEXTS=4
run ctx6 MAMRI_BENCH_CONTEXTS=6
run ctx12 MAMRI_BENCH_CONTEXTS=12
run ctx16 MAMRI_BENCH_CONTEXTS=16
run chains1 MAMRI_HBM_CHAINS=1
run chains3 MAMRI_HBM_CHAINS=3
run scan74 MAMRI_SCAN_CTAS=74
run scan296 MAMRI_SCAN_CTAS=296
run runctas37 MAMRI_RUN_CTAS=37
run runctas74 MAMRI_RUN_CTAS=74
run runctas296 MAMRI_RUN_CTAS=296
run slice128 MAMRI_SLICE_THREADS=128
run slice512 MAMRI_SLICE_THREADS=512
run close1 MAMRI_CLOSE_CTAS_PER_SM=1
run unfused MAMRI_CLOSE_FUSED=0
run thr8 MAMRI_THR_CTAS_PER_SM=8
run thr16 MAMRI_THR_CTAS_PER_SM=16
run mat4 MAMRI_MAT_CTAS_PER_SM=4
run mat8 MAMRI_MAT_CTAS_PER_SM=8
run nopdl MAMRI_PDL=0
run cluster16 MAMRI_LABEL_CLUSTER=16
MAMRI_WAVE_TRACE=1 python tools/batch_only.py > $O/wave_trace.log 2>&1
tail -12 $O/wave_trace.log
