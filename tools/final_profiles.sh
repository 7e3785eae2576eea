#!/bin/bash
# The artefact set profiles/README.md quotes, in one call (run under gpurun; everything lands in gpurun_out/<tag>_*).
# usage: tools/final_profiles.sh <tag> [tests]      e.g. tools/final_profiles.sh r2final tests
cd "$(dirname "$0")/.."
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
if [ -n "$2" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -4; fi
# 1. the bench line and the reference arm (what the driver runs)
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_ref.err
# 2. ncu launch list of the same command without the host legs (per-launch times: cold, serialised -> shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-detect > $O/${TAG}_launches.log 2>&1
# 3. one --set full capture of each main kernel (131 072 distinct reads, 1.97 Gbases per launch)
ncu --set full --clock-control none --import-source on -k regex:"k_cycle_stats|k_scan_jit|k_trim|k_final|k_read_qual|k_kmer_fix" \
    -s 21 -c 7 -f -o $O/prof_${TAG} python bench.py --steps 1 --warmup 3 --reads 131072 --no-e2e --no-parity --no-cpu-baseline --no-detect \
    > $O/ncu_${TAG}.log 2>&1
tail -2 $O/ncu_${TAG}.log
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["parity_checked"], d["e2e"]["value"], d["e2e"]["roofline"]["frac"], d["cpu_baseline"]["value"],
      d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("adapter_quality_kernel"))
print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
print(open("$O/${TAG}_bench_reference.json").read()[:300])
PY
