#!/bin/bash
# Short form of round_check.sh for the last GPU minutes of a round: the same commands minus the reference arm and the
# 26-connectivity C4 capture (three full C4/C2 reports together exceed gpurun's 64 MiB return limit).
set -u
T=${1:-r02d}
O=gpurun_out/$T
mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 120 python tools/profile_one.py --scans 3 > $O/plain.log 2>&1; echo "plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2.csv \
    python tools/profile_one.py --scans 3 > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_threshold|k_close|k_morph|k_runs|k_union|k_flatten|k_select|k_stats|k_materialise' -s 20 -c 10 -o $O/full_c2 -f \
    python tools/profile_one.py --scans 3 > $O/ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
for c in c1 c2 c3 c4; do timeout 120 python tools/ktrace.py --config $c --reps 20 > $O/ktrace_$c.txt 2>&1; done
timeout 120 python tools/ktrace.py --config c4 --conn 26 --reps 10 > $O/ktrace_c4_conn26.txt 2>&1
for c in c1 c2 c3 c4; do timeout 120 python tools/serial_latency.py --config $c --reps 30; done > $O/serial.txt 2>&1
timeout 120 python tools/serial_latency.py --config c4 --conn 26 --reps 30 >> $O/serial.txt 2>&1
cat $O/serial.txt
timeout 400 ncu --set full --clock-control none -k regex:'k_threshold|k_close|k_morph|k_runs|k_union|k_flatten|k_select|k_stats|k_materialise' -s 12 -c 12 -o $O/full_c4 -f \
    python tools/profile_one.py --config c4 --scans 2 > $O/ncu_full_c4.log 2>&1; echo "ncu full c4 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 900 --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-c4 --c3-scans 8 > $O/ncu_bench.log 2>&1; echo "ncu bench list rc=$?"
