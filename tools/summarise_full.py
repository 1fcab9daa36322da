"""Summarises an `ncu --set full` report (already exported with `ncu -i X.ncu-rep --page raw --csv`) into the
few metrics profiles/README.md discusses.  usage: summarise_full.py raw.csv > summary.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("##", r[idx["Kernel Name"]][:90])
    for w in want:
        if w in idx:
            print(f"  {w:62s} {r[idx[w]]} {units[idx[w]]}")
    print()
