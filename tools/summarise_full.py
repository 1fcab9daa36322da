"""Summarises an `ncu --set full` report into one table row per launch: the few metrics profiles/README.md discusses.
usage: summarise_full.py report.ncu-rep > table.txt   (runs `ncu -i ... --page raw --csv` itself)"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("us", "gpu__time_duration.sum", 1.0), ("grid", "launch__grid_size", 1), ("blk", "launch__block_size", 1),
        ("regs", "launch__registers_per_thread", 1), ("DRAM rd MB", "dram__bytes_read.sum", 1.0), ("DRAM wr MB", "dram__bytes_write.sum", 1.0),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0), ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0), ("Minst", "smsp__inst_executed.sum", 1e-6),
        ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
        ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0),
        ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0)]
units = rows[1]
print(f"{'kernel':34s}" + "".join(f"{c[0]:>14s}" for c in cols))
tot = 0.0
for r in rows[2:]:
    name = r[idx["Kernel Name"]].replace("void ", "")[:33]
    out = f"{name:34s}"
    for label, key, scale in cols:
        if key not in idx:
            out += f"{'-':>14s}"
            continue
        v = r[idx[key]].replace(",", "")
        try:
            x = float(v)
            u = units[idx[key]]
            if label.endswith("MB") and u == "byte": x /= 1e6
            if label.endswith("MB") and u == "Kbyte": x /= 1e3
            if label.endswith("MB") and u == "Gbyte": x *= 1e3
            if label == "us" and u == "ns": x /= 1e3
            if label == "us" and u == "ms": x *= 1e3
            if label == "us": tot += x
            out += f"{x * scale if isinstance(scale, float) and label == 'Minst' else x:14.2f}"
        except ValueError:
            out += f"{v:>14s}"
    print(out)
print(f"sum of the launch durations: {tot:.1f} us (cold cache, serialised -- compare shares, not absolutes)")
