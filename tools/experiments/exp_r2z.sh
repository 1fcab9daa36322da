#!/bin/bash
# nbr_range: the first entries of both searches loaded together
set -u
timeout 90 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q --deselect tests/test_gpu_parity.py::test_split_statistics_path_on_small_scans 2>&1 | tail -2
for c in c1 c2 c4; do timeout 30 python tools/serial_latency.py --config $c --reps 30 2>&1 | sed 's/.*bare C ABI/  '$c' bare/'; done
