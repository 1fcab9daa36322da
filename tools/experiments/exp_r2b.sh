#!/bin/bash
set -u
O=gpurun_out/r2b
mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
MAMRI_LABEL_CLUSTER=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest_scalable.log 2>&1; echo "pytest scalable rc=$?"; tail -3 $O/pytest_scalable.log
T="python tools/ktrace.py --reps 20"
$T --config c2 > $O/kt_c2.log 2>&1; echo "kt c2 rc=$?"
$T --config c1 > $O/kt_c1.log 2>&1
$T --config c3 > $O/kt_c3.log 2>&1
MAMRI_LABEL_CLUSTER=0 $T --config c2 > $O/kt_c2_scalable.log 2>&1
MAMRI_LABEL_CLUSTER=8 $T --config c2 > $O/kt_c2_cl8.log 2>&1
MAMRI_CLUSTER_FENCE=0 $T --config c2 > $O/kt_c2_nofence.log 2>&1
MAMRI_CLOSE_SMEM_KB=200 MAMRI_CLOSE_TY=32 MAMRI_CLOSE_TZ=16 MAMRI_CLOSE_ZSD=2 MAMRI_CLOSE_ZSE=2 $T --config c2 > $O/kt_c2_close32.log 2>&1
MAMRI_CLOSE_TZ=8 $T --config c2 > $O/kt_c2_close_tz8.log 2>&1
$T --config c4 --reps 5 > $O/kt_c4.log 2>&1
MAMRI_CLOSE_SMEM_KB=200 $T --config c4 --reps 5 > $O/kt_c4_smem200.log 2>&1
MAMRI_CLOSE_FUSED=0 $T --config c4 --reps 5 > $O/kt_c4_unfused.log 2>&1
$T --config c4 --conn 26 --reps 5 > $O/kt_c4_26.log 2>&1
for c in c1 c2 c3 c4; do python tools/serial_latency.py --config $c --reps 30 > $O/serial_$c.log 2>&1; done
python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
MAMRI_LABEL_CLUSTER=0 python bench.py --no-cpu-baseline > $O/bench_scalable.json 2>> $O/bench.err
cat $O/serial_*.log
grep -h "event-timed" $O/kt_*.log
