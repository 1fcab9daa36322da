"""Copies the outputs of `tools/round_check.sh TAG` from gpurun_out/TAG/ into profiles/ (tracked), writes the summaries of the
ncu captures and profiles/traffic.json (read by bench.py for `roofline.traffic`).  Missing pieces are skipped with a note, so
a round check that was cut short still publishes what it has.       usage: python tools/publish_profiles.py TAG"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")


def have(name):
    p = os.path.join(G, name)
    ok = os.path.exists(p) and os.path.getsize(p) > 0
    if not ok:
        print(f"-- {name}: absent, skipped")
    return ok


def run(args, out, header=""):
    with open(out, "w") as f:
        f.write(header + subprocess.run(args, check=True, capture_output=True, text=True).stdout)


def last_json(path):
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


kps = 10
if have("bench.json"):
    d = last_json(f"{G}/bench.json")
    if d:
        kps = int(d.get("kernels_per_scan", kps))
        shutil.copy(f"{G}/bench.json", f"{P}/{tag}_bench_n1.json")
if have("bench_ref.json"):
    shutil.copy(f"{G}/bench_ref.json", f"{P}/{tag}_bench_reference.json")
if have("launches_c2.csv"):
    shutil.copy(f"{G}/launches_c2.csv", f"{P}/{tag}_launches_c2.csv")
    run([sys.executable, f"{ROOT}/tools/parse_launches.py", f"{G}/launches_c2.csv", str(kps)], f"{P}/{tag}_launches_c2_summary.txt")
if have("launches_bench.csv"):
    shutil.copy(f"{G}/launches_bench.csv", f"{P}/{tag}_launches_bench.csv")
    try:
        run([sys.executable, f"{ROOT}/tools/summarise_bench_launches.py", f"{G}/launches_bench.csv", "8", str(kps)],
            f"{P}/{tag}_launches_bench_summary.txt")
    except subprocess.CalledProcessError as e:
        print("-- bench launch summary failed:", e.stderr[-300:])
for rep, what in (("full_c2", "C2 (512x512x256 u16), third scan of three"), ("full_c4", "C4 (1024x1024x512 u16, 6-connectivity), second scan of two"),
                  ("full_c4_conn26", "C4 (1024x1024x512 u16, 26-connectivity), second scan of two")):
    if have(f"{rep}.ncu-rep"):
        run([sys.executable, f"{ROOT}/tools/summarise_full.py", f"{G}/{rep}.ncu-rep"], f"{P}/{tag}_ncu_{rep}_all_kernels.txt",
            f"ncu --set full --clock-control none, every kernel of one scan, {what}; one row per launch\n\n")
for f in sorted(os.listdir(G)) if os.path.isdir(G) else []:
    if f.startswith("ktrace_") or f in ("serial.txt", "pytest.log", "smoke.log"):
        if have(f):
            shutil.copy(f"{G}/{f}", f"{P}/{tag}_{f}")

# per-launch DRAM traffic of the two streaming kernels -> traffic.json
tab = f"{P}/{tag}_ncu_full_c2_all_kernels.txt"
if os.path.exists(tab):
    lines = open(tab).read().splitlines()
    hdr = [l for l in lines if l.startswith("kernel")][0]
    cols = re.split(r"\s{2,}", hdr.strip())

    def row(prefix):
        for l in lines:
            if l.startswith(prefix):
                vals = l[34:].split()
                return dict(zip(cols[1:], vals))
        return None

    t, m = row("k_threshold"), row("k_materialise")
    if t and m:
        traffic = {"threshold_pack": (float(t["DRAM rd MB"]) + float(t["DRAM wr MB"])) * 1e6,
                   "materialise": (float(m["DRAM rd MB"]) + float(m["DRAM wr MB"])) * 1e6,
                   "_source": f"profiles/{tag}_ncu_full_c2_all_kernels.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch, C2 scan "
                              "(512x512x256 u16), ncu --set full"}
        json.dump(traffic, open(f"{P}/traffic.json", "w"), indent=1)
        print(traffic, "materialise us", m["us"], "threshold us", t["us"])
d = last_json(f"{P}/{tag}_bench_n1.json") if os.path.exists(f"{P}/{tag}_bench_n1.json") else None
if d:
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "kernel frac", d["roofline"]["frac"],
          "launch_ms", d["roofline"]["launch_ms"], "pipeline", d["roofline"]["pipeline"], "cpu", d["cpu_baseline"]["value"], d["stages_ms"])
