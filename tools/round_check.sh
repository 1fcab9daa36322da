#!/bin/bash
# Every command runs under its own `timeout`: a kernel that hangs costs minutes, not the rest of the GPU budget.
# Full GPU check of the current tree: smoke, GPU tests, both bench arms, in-pipeline timelines, ncu launch list + full
# capture of every kernel of a C2 scan and of a C4 scan.  Each ncu pass runs only after the same command exited 0 without ncu.
set -u
T=${1:-r02}
O=gpurun_out/$T
mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench.json 2>> $O/bench.err; echo "bench rc=$?"
timeout 180 python tools/profile_one.py --scans 3 > $O/plain.log 2>&1; echo "plain rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2.csv \
    python tools/profile_one.py --scans 3 > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 900 --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-c4 --c3-scans 8 > $O/ncu_bench.log 2>&1; echo "ncu bench list rc=$?"
# every kernel of the third C2 scan (10 launches), full set
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_threshold|k_close|k_morph|k_runs|k_union|k_flatten|k_select|k_stats|k_materialise' -s 20 -c 10 -o $O/full_c2 -f \
    python tools/profile_one.py --scans 3 > $O/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
for c in c1 c2 c3 c4; do timeout 180 python tools/ktrace.py --config $c --reps 20 > $O/ktrace_$c.txt 2>&1; done
timeout 180 python tools/ktrace.py --config c4 --conn 26 --reps 10 > $O/ktrace_c4_conn26.txt 2>&1
for c in c1 c2 c3 c4; do timeout 180 python tools/serial_latency.py --config $c --reps 30; done > $O/serial.txt 2>&1
timeout 180 python tools/profile_one.py --config c4 --scans 2 > $O/plain_c4.log 2>&1; echo "plain c4 rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:'k_threshold|k_close|k_morph|k_runs|k_union|k_flatten|k_select|k_stats|k_materialise' -s 12 -c 12 -o $O/full_c4 -f \
    python tools/profile_one.py --config c4 --scans 2 > $O/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:'k_threshold|k_close|k_morph|k_runs|k_union|k_flatten|k_select|k_stats|k_materialise' -s 12 -c 12 -o $O/full_c4_conn26 -f \
    python tools/profile_one.py --config c4 --scans 2 --conn 26 > $O/ncu_full_c4_26.log 2>&1; echo "ncu full c4/26 rc=$?"
cat $O/serial.txt

# NOTE: the three .ncu-rep files together are ~64 MiB, gpurun's limit for what it copies back: if the call reports
# "gpurun_out/ not copied back", use tools/round_final.sh (no C4/26 capture) and empty gpurun_out/ here first.
