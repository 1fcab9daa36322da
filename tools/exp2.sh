#!/bin/bash
out=gpurun_out/exp2.log; : > $out
run() { echo "== $*" >> $out; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>$out | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'Gvox/s', 'e2e', round(d['e2e']['value'],1), d['stages_ms'], d['roofline']['frac'])" >> $out; }
run MAMRI_MORPH_PLANES=1
run MAMRI_MORPH_PLANES=0
run MAMRI_BENCH_CONTEXTS=2
run MAMRI_BENCH_CONTEXTS=3
run MAMRI_BENCH_CONTEXTS=6
run MAMRI_BENCH_CONTEXTS=8
run MAMRI_BENCH_CONTEXTS=8 MAMRI_PRIO_SMALL=0 MAMRI_PRIO_BIG=0 MAMRI_PDL=0
run MAMRI_BENCH_CONTEXTS=8 MAMRI_MAT_CTAS_PER_SM=64 MAMRI_THR_CTAS_PER_SM=64
for p in 0 1; do echo "== serial PLANES=$p" >> $out; MAMRI_MORPH_PLANES=$p python tools/serial_latency.py >> $out 2>&1; done
python tools/profile_one.py --scans 3 >> $out 2>&1
python tools/host_overhead.py >> $out 2>&1
cat $out
MAMRI_PDL=1 python tools/serial_latency.py --config c4 --reps 10 >> gpurun_out/exp2.log 2>&1; python tools/profile_one.py --config c4 --scans 2 >> gpurun_out/exp2.log 2>&1; tail -30 gpurun_out/exp2.log
