"""Prints the per-kernel durations of the LAST scan in an ncu --metrics gpu__time_duration.sum CSV log."""
import csv, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]) for r in csv.DictReader(lines)]
tot = sum(v for _, v, _, _ in rows[-n:])
for name, v, g, b in rows[-n:]:
    print(f"{v/1000:9.2f} us {100*v/tot:5.1f}%  {g:>14} {b:>12}  {name[:64]}")
print(f"sum {tot/1000:.1f} us over {n} launches")
