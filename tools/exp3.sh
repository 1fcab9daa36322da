#!/bin/bash
out=gpurun_out/exp3.log; : > $out
run() { echo "== $*" >> $out; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>$out | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'Gvox/s', 'e2e', round(d['e2e']['value'],1), d['stages_ms'], d['roofline']['frac'])" >> $out; }
run MAMRI_BENCH_CONTEXTS=8
run MAMRI_BENCH_CONTEXTS=8 MAMRI_TILE_TZ=8
run MAMRI_BENCH_CONTEXTS=8 MAMRI_MORPH_PLANES=1
for p in 16 8 4; do echo "== serial TZ=$p" >> $out; MAMRI_TILE_TZ=$p python tools/serial_latency.py >> $out 2>&1; done
python tools/profile_one.py --scans 3 >> $out 2>&1
python tools/profile_one.py --config c4 --scans 2 >> $out 2>&1
cat $out
