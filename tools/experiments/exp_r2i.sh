#!/bin/bash
set -u
O=gpurun_out/r2i
mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for sp in 0 1 2; do
  echo "== MAMRI_MAT_SPLIT=$sp"
  for c in c1 c2 c3 c4; do MAMRI_MAT_SPLIT=$sp python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
done
echo "== MAMRI_MAT_SPLIT=3 c4"; MAMRI_MAT_SPLIT=3 python tools/serial_latency.py --config c4 --reps 30 2>&1 | sed 's/.*bare C ABI/  c4 bare/'
echo "== SCAN_CTAS"
for n in 74 296; do for c in c2 c4; do MAMRI_SCAN_CTAS=$n python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' ctas='$n' bare/'; done; done
for sp in 0 2; do MAMRI_MAT_SPLIT=$sp python tools/ktrace.py --config c2 --reps 20 > $O/kt_c2_split$sp.log 2>&1; done
python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4.log 2>&1
MAMRI_MAT_SPLIT=3 python tools/ktrace.py --config c4 --reps 5 > $O/kt_c4_split3.log 2>&1
python bench.py --no-cpu-baseline --skip-c4 --c3-scans 8 --steps 60 > $O/bench.json 2>$O/bench.err; tail -c 600 $O/bench.json | head -c 300; echo
ncu --set full --clock-control none -k regex:'k_runs|k_stats|k_materialise|k_union|k_flatten' -s 7 -c 7 -o $O/full_c4 -f \
    python tools/profile_one.py --config c4 --scans 2 > $O/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
