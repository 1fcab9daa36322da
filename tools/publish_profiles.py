"""Copies the outputs of `tools/round_check.sh TAG` (+ tools/bench_configs.py) from gpurun_out/ into profiles/ and writes
the summaries and profiles/traffic.json.  usage: python tools/publish_profiles.py TAG"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(args, out):
    with open(out, "w") as f:
        f.write(subprocess.run(args, check=True, capture_output=True, text=True).stdout)


shutil.copy(f"{G}/launches_{tag}.csv", f"{P}/{tag}_launches_c2.csv")
run([sys.executable, f"{ROOT}/tools/parse_launches.py", f"{G}/launches_{tag}.csv", "12"], f"{P}/{tag}_launches_c2_summary.txt")
shutil.copy(f"{G}/launches_bench_{tag}.csv", f"{P}/{tag}_launches_bench.csv")
run([sys.executable, f"{ROOT}/tools/summarise_bench_launches.py", f"{G}/launches_bench_{tag}.csv"], f"{P}/{tag}_launches_bench_summary.txt")
raw = f"/tmp/full_{tag}.csv"
with open(raw, "w") as f:
    f.write(subprocess.run(["ncu", "-i", f"{G}/full_{tag}.ncu-rep", "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout)
summ = subprocess.run([sys.executable, f"{ROOT}/tools/summarise_full.py", raw], check=True, capture_output=True, text=True).stdout
with open(f"{P}/{tag}_ncu_full_c2_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on, C2 (512x512x256 u16), 2 scans, kernels k_threshold_pack* and "
            "k_materialise*; per launch\n\n" + summ)
for src, dst in ((f"bench_{tag}.json", f"{tag}_bench_n1.json"), (f"bench_ref_{tag}.json", f"{tag}_bench_reference.json"),
                 (f"configs_{tag}.jsonl", f"{tag}_configs.jsonl")):
    if os.path.exists(f"{G}/{src}"):
        shutil.copy(f"{G}/{src}", f"{P}/{dst}")
blocks = summ.split("## ")[1:]


def val(b, k):
    return float(re.search(k + r"\s+([0-9.]+)", b).group(1))


t = [b for b in blocks if "k_threshold" in b][0]
m = [b for b in blocks if "k_materialise" in b][0]
traffic = {"threshold_pack": (val(t, "dram__bytes_read.sum") + val(t, "dram__bytes_write.sum")) * 1e6,
           "materialise": (val(m, "dram__bytes_read.sum") + val(m, "dram__bytes_write.sum")) * 1e6,
           "_source": f"profiles/{tag}_ncu_full_c2_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch, C2 scan "
                      "(512x512x256 u16), ncu --set full"}
json.dump(traffic, open(f"{P}/traffic.json", "w"), indent=1)
print(traffic, "materialise us", val(m, "gpu__time_duration.sum"), "threshold us", val(t, "gpu__time_duration.sum"))
d = json.loads(open(f"{G}/bench_{tag}.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "kernel frac", d["roofline"]["frac"],
      "launch_ms", d["roofline"]["launch_ms"], "pipeline", d["roofline"]["pipeline"], "cpu", d["cpu_baseline"]["value"], d["stages_ms"])
