#!/bin/bash
# The other BASELINE configurations with the current build (run under gpurun): one bench line each, parity-checked
# against the oracle on its own input, into gpurun_out/<tag>_bench_<workload>.json.
# usage: tools/other_configs.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-cfg}
O=gpurun_out
mkdir -p $O
run() { python bench.py --workload "$1" ${2:+--reads $2} --steps ${3:-5} --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_$1.json 2> $O/${TAG}_bench_$1.err; }
run c3 131072 3
run c4 1000000
run c5-1k
run c5-50k
run c5-500k
python - <<PY
import json
for w in ("c3", "c4", "c5-1k", "c5-50k", "c5-500k"):
    try:
        d = json.load(open("$O/${TAG}_bench_%s.json" % w))
        print(w, d["value"], d["ms_per_step"], d.get("parity_checked"), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
    except Exception as e:
        print(w, "failed:", e)
PY
