#!/bin/bash
# k_stats: run extents loaded ahead; k_flatten_rank: joint root walks + early publish; k_union_z: both finds together;
# k_union_slices: flags for the runs that touch several runs above
set -u
O=gpurun_out/r2t
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for c in c1 c2 c3 c4; do timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
timeout 120 python tools/serial_latency.py --config c4 --conn 26 --reps 30 2>&1 | sed 's/.*bare C ABI/  c4-26 bare/'
timeout 120 python tools/ktrace.py --config c4 --reps 10 > $O/kt_c4.log 2>&1; echo "== c4"; cat $O/kt_c4.log
timeout 120 python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2.log 2>&1; echo "== c2"; cat $O/kt_c2.log | grep -vE "close\.|uslice\."
timeout 120 python tools/batch_only.py
