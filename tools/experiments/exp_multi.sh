#!/bin/bash
# N-GPU runs of the bench line (as the driver launches it), the no-gather arm and the copy-ceiling probe.
set -u
N=${1:-2}
O=gpurun_out/multi
mkdir -p $O
P=$((29500 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P"
$TR bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n=$N rc=$?"
MAMRI_BENCH_NO_GATHER=1 $TR bench.py --gpus $N --no-cpu-baseline --skip-c4 --c3-scans 16 > $O/bench_nogather_n$N.json 2>> $O/bench_n$N.err; echo "nogather rc=$?"
MAMRI_BENCH_BODY_U8=1 $TR bench.py --gpus $N --no-cpu-baseline --skip-c4 --c3-scans 16 > $O/bench_u8_n$N.json 2>> $O/bench_n$N.err; echo "u8 rc=$?"
$TR tools/pcie_probe.py > $O/pcie_n$N.json 2>> $O/bench_n$N.err; echo "pcie rc=$?"
python bench.py --impl reference --gpus $N > $O/bench_ref_n$N.json 2>> $O/bench_n$N.err
tail -3 $O/bench_n$N.err
python - $O $N <<'PY'
import json, sys
O, N = sys.argv[1], sys.argv[2]
for name in (f"bench_n{N}", f"bench_nogather_n{N}", f"bench_u8_n{N}"):
    try:
        d = json.loads(open(f"{O}/{name}.json").read().strip().splitlines()[-1])
        print(name, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "ranks", d["rank_ms_per_step"], "e2e", round(d["e2e"]["value"], 2),
              "ceil", round(d["e2e"]["copy_ceiling"]["value"], 2), "frac", round(d["e2e"]["frac_of_copy_ceiling"], 3),
              "C3", {k: d["configs"]["C3"][k] for k in ("ms_per_batch", "scans_per_s", "gathered_equals_single_gpu")},
              "C5", {k: (round(v["ms"], 4), v.get("sharded_equals_single_gpu")) for k, v in d["configs"]["C5"].items() if k.startswith("path")},
              "gathered_host", d.get("gathered_tables_equal_host_packed"))
    except Exception as e:
        print(name, "ERR", e)
print(open(f"{O}/pcie_n{N}.json").read())
PY
