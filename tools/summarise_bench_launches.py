"""Summarises the ncu `--metrics gpu__time_duration.sum` launch list of `python bench.py --steps 2 --warmup 3` (the
pass tools/round_check.sh records): the two timed steps, per kernel, averaged per scan.
usage: summarise_bench_launches.py launches_bench.csv [scans_per_step=8] [kernels_per_scan=12] [warm_steps=4]"""
import collections
import csv
import sys

path = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
KPS = int(sys.argv[3]) if len(sys.argv) > 3 else 12
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
names = [r["Kernel Name"] for r in rows]
first = [i for i, n in enumerate(names) if "k_threshold" in n][0]       # after the phantom generator's kernels
per_step = S * KPS
start = first + warm * per_step
seg = rows[start:start + 2 * per_step]
agg = collections.OrderedDict()
for r in seg:
    n = r["Kernel Name"].split("(")[0][:56]
    a = agg.setdefault(n, [0.0, 0])
    a[0] += float(r["Metric Value"].replace(",", ""))
    a[1] += 1
tot = sum(v[0] for v in agg.values())
n_scans = 2 * S
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print(f"{len(rows)} launches captured; below: the 2 timed steps ({n_scans} scans, launches {start}..{start + len(seg) - 1}: after {first} set-up")
print(f"kernels and {warm} warm-up steps of {per_step} launches), per kernel, averaged per scan.  Kernels inside the captured wave graphs")
print("are profiled node by node (serialised, cold cache): compare shares with bench.py's stages_ms, not absolutes.\n")
for n, (v, c) in agg.items():
    print(f"{v / n_scans / 1000:8.2f} us/scan {100 * v / tot:5.1f}%  x{c // n_scans} per scan  {n}")
print(f"sum {tot / n_scans / 1000:.1f} us per scan over {sum(c for _, c in agg.values()) // n_scans} launches")
