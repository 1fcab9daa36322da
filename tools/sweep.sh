#!/bin/bash
# sweep of grid sizes for the two DRAM-bound kernels (persistent-ish grids leave SM slots for other streams' small kernels)
for cfg in "16 32 4" "8 32 4" "4 32 4" "2 32 4" "4 8 4" "4 4 4" "2 4 4" "2 4 6" "4 8 6"; do
  set -- $cfg
  MAMRI_MAT_CTAS_PER_SM=$1 MAMRI_THR_CTAS_PER_SM=$2 MAMRI_BENCH_CONTEXTS=$3 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mat/thr/ctx', '$1', '$2', '$3', round(d['value'],1), 'Gvox/s', d['stages_ms'])"
done
