#!/bin/bash
set -u
O=gpurun_out/r2j
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for st in 1 0; do for ilp in 0 1; do
  echo "== RUNS_STAGE=$st MAT_ILP=$ilp"
  for c in c2 c4; do MAMRI_RUNS_STAGE=$st MAMRI_MAT_ILP=$ilp timeout 120 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
done; done
for st in 1 0; do
  MAMRI_RUNS_STAGE=$st timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4_stage$st.log 2>&1
  echo "== c4 stage=$st"; grep -E "^ +(runs_scan|runs.lookback|runs.lastCTA|union_slices|materialise|end) " $O/kt_c4_stage$st.log
  MAMRI_RUNS_STAGE=$st timeout 120 python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2_stage$st.log 2>&1
  echo "== c2 stage=$st"; grep -E "^ +(runs_scan|runs.lookback|runs.lastCTA|union_slices|materialise|end) " $O/kt_c2_stage$st.log
done
MAMRI_MAT_ILP=1 timeout 120 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4_ilp.log 2>&1
echo "== c4 ilp"; grep -E "^ +(materialise|stats.finalise|final|end) " $O/kt_c4_ilp.log
echo "== threshold CTAs per SM (c2)"
for n in 6 12 18 24 32; do
  MAMRI_THR_CTAS_PER_SM=$n timeout 120 python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2_thr$n.log 2>&1
  echo "  per_sm=$n: $(grep -E '^ +threshold.lastCTA' $O/kt_c2_thr$n.log | awk '{print $2}') us; $(MAMRI_THR_CTAS_PER_SM=$n timeout 120 python tools/serial_latency.py --config c2 --reps 30 2>&1 | sed 's/.*bare C ABI//')"
done
