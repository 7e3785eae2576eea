"""Development aid (CPU): adapter auto-detection — fastplong_b200/evaluator.py (the host half of
Evaluator::evalAdapterAndReadNum, src/evaluator.cpp:105-265) on the numpy ten-mer tables, with the table half taken once from the C ABI (fpl_eval_pick_adapter) and once from its
Python twin, against what the unmodified
reference BINARY detects on the same FASTQ (its JSON's adapter_cutting.read_start_adapter / read_end_adapter).
Random adapters (plain, low-complexity, repetitive, G-rich, short), random presence rates, read counts around the
100-read rule, read lengths around the 128-base evaluation window, --trim_tail values.
usage: python tools/fuzz_evaluator_vs_binary.py <seed> <seconds>"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastplong_b200 import evaluator, synth  # noqa: E402
from oracle_lib import REF_BIN, kmer10_tables  # noqa: E402


def random_adapter(rng):
    kind = rng.integers(0, 6)
    n = int(rng.integers(12, 50))
    if kind == 0:      # plain
        return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
    if kind == 1:      # two-letter alphabet: the low-complexity rules of getTopKey
        ab = rng.choice(4, size=2, replace=False)
        return "".join("ACGT"[ab[i]] for i in rng.integers(0, 2, size=n))
    if kind == 2:      # short period
        unit = "".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(2, 8))))
        return (unit * 30)[:n]
    if kind == 3:      # G/C rich
        return "".join("GGGCGCAT"[i] for i in rng.integers(0, 8, size=n))
    if kind == 4:      # starts with a G run
        return "G" * int(rng.integers(4, 9)) + "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(6, 17))))   # at / below the length-16 rule


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = bad = 0
    found = 0
    while time.time() - t0 < budget:
        a_s, a_e = random_adapter(rng), random_adapter(rng)
        kw = dict(adapter_start=a_s, adapter_end=a_e, p_start=float(rng.choice([0.0, 0.05, 0.3, 0.6, 0.9])),
                  p_end=float(rng.choice([0.0, 0.05, 0.3, 0.6, 0.9])), p_chimera=float(rng.choice([0.0, 0.05])),
                  p_polya=float(rng.choice([0.0, 0.0, 0.3])), min_len=int(rng.choice([20, 100, 200])))
        n_reads = int(rng.choice([90, 101, 300, 800, 2000]))
        mean = int(rng.choice([60, 150, 400, 1500]))
        trim_tail = int(rng.choice([0, 0, 0, 3, 10]))
        batch = synth.ont_like(n_reads, mean, int(rng.integers(1 << 30)), **kw)
        with tempfile.TemporaryDirectory() as d:
            fq = os.path.join(d, "in.fq")
            synth.to_fastq(batch, fq)
            cmd = [REF_BIN, "-i", fq, "-o", d + "/o.fq", "-j", d + "/j.json", "-h", d + "/h.html", "-w", "2"]
            if trim_tail:
                cmd += ["-t", str(trim_tail)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                print("reference refused:", r.stderr[-200:])
                continue
            ac = json.load(open(d + "/j.json")).get("adapter_cutting", {})
        ref = (ac.get("read_start_adapter"), ac.get("read_end_adapter"))
        found += sum(s != "unspecified" for s in ref)
        for pick in ("abi", "python"):      # the C ABI's fpl_eval_pick_adapter and evaluator.detect_one
            got = evaluator.detect_adapters(batch, trim_tail=trim_tail, kmers=kmer10_tables, pick=pick)
            got = tuple("unspecified" if s == "auto" else s for s in got)     # src/options.cpp:247-259
            if got != ref:
                bad += 1
                print("MISMATCH case", n, pick, "reads", n_reads, "mean", mean, "trim_tail", trim_tail, kw, "\n  ours", got, "\n  ref ", ref)
        if bad > 5:
            break
        n += 1
    print("cases", n, "mismatches", bad, "(adapters the reference detected:", found, ")")


if __name__ == "__main__":
    main()
