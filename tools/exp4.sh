#!/bin/bash
out=gpurun_out/exp4.log; : > $out
run() { echo "== $*" >> $out; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>$out | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'Gvox/s', 'e2e', round(d['e2e']['value'],1), d['stages_ms'], d['roofline']['frac'])" >> $out; }
run MAMRI_BENCH_CONTEXTS=4
run MAMRI_BENCH_CONTEXTS=8
run MAMRI_BENCH_CONTEXTS=12
python tools/serial_latency.py >> $out 2>&1
python tools/profile_one.py --scans 3 >> $out 2>&1
python tools/profile_one.py --config c4 --scans 2 >> $out 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "small_phantoms or all_radii" >> $out 2>&1; echo "memcheck rc=$?" >> $out
cat $out
