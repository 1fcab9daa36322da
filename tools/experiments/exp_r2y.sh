#!/bin/bash
# last-minute sweep: streaming kernels with a smaller footprint (fewer resident CTAs), so that the labelling kernels of the
# other scans of a wave find free registers / thread slots
set -u
run() { echo -n "$* : "; env "$@" STEPS=30 timeout 60 python tools/pipe_only.py 2>&1 | tail -1; }
run A=1
run MAMRI_THR_CTAS_PER_SM=2
run MAMRI_THR_CTAS_PER_SM=3
run MAMRI_THR_CTAS_PER_SM=4
run MAMRI_MAT_CTAS_PER_SM=5
run MAMRI_THR_CTAS_PER_SM=3 MAMRI_MAT_CTAS_PER_SM=6
