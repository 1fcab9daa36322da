#!/bin/bash
# A/B of launch priorities / programmatic dependent launch / context count (writes gpurun_out/exp_overlap.log)
out=gpurun_out/exp_overlap.log; : > $out
run() { echo "== $*" >> $out; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>$out | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'Gvox/s', 'e2e', round(d['e2e']['value'],1), d['stages_ms'], d['roofline']['frac'])" >> $out; }
run MAMRI_PRIO_SMALL=0 MAMRI_PRIO_BIG=0 MAMRI_PDL=0
run MAMRI_PDL=0
run MAMRI_PRIO_SMALL=0 MAMRI_PRIO_BIG=0
run MAMRI_PDL=1
run MAMRI_BENCH_CONTEXTS=2
run MAMRI_BENCH_CONTEXTS=6
run MAMRI_BENCH_CONTEXTS=8
run MAMRI_BENCH_CONTEXTS=8 MAMRI_MAT_CTAS_PER_SM=8 MAMRI_THR_CTAS_PER_SM=16
run MAMRI_BENCH_CONTEXTS=8 MAMRI_MAT_CTAS_PER_SM=4 MAMRI_THR_CTAS_PER_SM=8
run MAMRI_BENCH_CONTEXTS=8 MAMRI_MAT_CTAS_PER_SM=64 MAMRI_THR_CTAS_PER_SM=64
for p in 0 1; do echo "== serial PDL=$p" >> $out; MAMRI_PDL=$p python tools/serial_latency.py >> $out 2>&1; done
MAMRI_PDL=1 python tools/serial_latency.py --config c4 --reps 10 >> $out 2>&1
cat $out
