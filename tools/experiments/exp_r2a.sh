#!/bin/bash
# Round-2 experiment A: correctness of the fused kernels + in-pipeline timelines + tunable sweeps.
set -u
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/smi.txt 2>&1
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
MAMRI_LABEL_CLUSTER=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest_scalable.log 2>&1; echo "pytest scalable rc=$?"; tail -3 $O/pytest_scalable.log
MAMRI_LABEL_CLUSTER=8 MAMRI_CLOSE_FUSED=0 MAMRI_THR_V8=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_c8.log 2>&1; echo "pytest cluster8/unfused rc=$?"; tail -3 $O/pytest_c8.log
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/memcheck.log
T="python tools/ktrace.py --reps 20"
$T --config c2 > $O/kt_c2.log 2>&1; echo "kt c2 rc=$?"
$T --config c1 > $O/kt_c1.log 2>&1
$T --config c3 > $O/kt_c3.log 2>&1
MAMRI_LABEL_CLUSTER=0 MAMRI_CLOSE_FUSED=0 MAMRI_THR_V8=0 $T --config c2 > $O/kt_c2_old.log 2>&1
MAMRI_LABEL_CLUSTER=0 $T --config c2 > $O/kt_c2_scalable.log 2>&1
MAMRI_LABEL_CLUSTER=8 $T --config c2 > $O/kt_c2_cl8.log 2>&1
MAMRI_CLUSTER_FENCE=0 $T --config c2 > $O/kt_c2_nofence.log 2>&1
MAMRI_THR_V8=0 $T --config c2 > $O/kt_c2_v4.log 2>&1
MAMRI_NO_GRAPH=1 $T --config c2 > $O/kt_c2_nograph.log 2>&1
MAMRI_PDL=0 $T --config c2 > $O/kt_c2_nopdl.log 2>&1
for v in "16 16 4 2 1 1" "16 16 4 4 1 1" "16 16 2 2 1 1" "16 8 4 2 1 1" "16 8 2 2 1 1" "8 8 4 2 1 1" "16 16 4 2 2 2" "16 16 2 2 2 2" "32 16 4 2 2 2" "32 16 4 4 1 1" "32 8 4 2 1 1"; do
  set -- $v
  MAMRI_CLOSE_SMEM_KB=200 MAMRI_CLOSE_TY=$1 MAMRI_CLOSE_TZ=$2 MAMRI_CLOSE_SYD=$3 MAMRI_CLOSE_SYE=$4 MAMRI_CLOSE_ZSD=$5 MAMRI_CLOSE_ZSE=$6 $T --config c2 --reps 10 > $O/kt_c2_close_$1_$2_$3_$4_$5_$6.log 2>&1
done
for p in 4 8 12 16 24; do MAMRI_THR_CTAS_PER_SM=$p $T --config c2 --reps 10 > $O/kt_c2_thr$p.log 2>&1; done
$T --config c4 --reps 5 > $O/kt_c4.log 2>&1
python tools/serial_latency.py --config c2 > $O/serial_c2.log 2>&1
python tools/serial_latency.py --config c1 > $O/serial_c1.log 2>&1
python tools/serial_latency.py --config c4 --reps 10 > $O/serial_c4.log 2>&1
python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
MAMRI_LIB= python bench.py --no-cpu-baseline --steps 50 > $O/bench2.json 2>> $O/bench.err
grep -h "event-timed" $O/kt_*.log | head -50
