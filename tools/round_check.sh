#!/bin/bash
# Full GPU check of the current tree: smoke, GPU tests, both bench arms, ncu launch list + full capture.
set -u
T=${1:-r01b}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$T.log
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_$T.err; echo "ref rc=$?"
python tools/profile_one.py --scans 3 > gpurun_out/plain_$T.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$T.csv \
    python tools/profile_one.py --scans 3 > gpurun_out/ncu_$T.log 2>&1; echo "ncu list rc=$?"
# the same pass over bench.py itself (kernels inside the captured wave graphs are profiled node by node)
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 900 --log-file gpurun_out/launches_bench_$T.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1; echo "ncu bench list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_materialise|k_threshold_pack' -c 4 \
    -o gpurun_out/full_$T -f python tools/profile_one.py --scans 2 > gpurun_out/ncu_full_$T.log 2>&1; echo "ncu full rc=$?"
