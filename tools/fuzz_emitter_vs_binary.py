"""Development aid (CPU): random option sets x seeded batches (tests/cases.py:random_case) — the host-side output rules
(fastplong_b200/hostside.py: names, split / break prefixes, masked bases, failed-out records) applied to the ORACLE's
records against the unmodified reference BINARY's --out / --failed_out files.
usage: python tools/fuzz_emitter_vs_binary.py <seed> <seconds>"""
import hashlib
import os
import random
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from fastplong_b200 import hostside, synth  # noqa: E402
from oracle_lib import REF_BIN, OracleEngine  # noqa: E402

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
t0 = time.time()
n = bad = refused = 0
while time.time() - t0 < budget:
    opt, batch, what = cases.random_case(rng)
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "in.fq")
        synth.to_fastq(batch, fq)
        extra = []
        if opt.adapter_fasta:
            fa = os.path.join(d, "a.fa")
            with open(fa, "w") as f:
                for i, s in enumerate(opt.adapter_fasta):
                    f.write(f">a{i:03d}\n{s}\n")
            extra = ["-a", fa]
        cmd = [REF_BIN, "-i", fq, "-o", d + "/o.fq", "--failed_out", d + "/f.fq", "-j", d + "/j.json", "-h", d + "/h.html",
               "-w", str(rng.choice([1, 2, 3]))] + opt.cli_flags() + extra
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:        # e.g. "This data contains both U and T" on the adversarial batches
            refused += 1
            continue
        ro, rf = open(d + "/o.fq", "rb").read(), open(d + "/f.fq", "rb").read()
    o = OracleEngine(opt)
    res = o.process(batch)
    names = hostside.default_names(batch)
    if opt.mask or opt.break_reads:
        out, failed = hostside.emit_fastq_ext(batch, names, res, o.segments(), o.mask_regions())
    else:
        out, failed = hostside.emit_fastq(batch, names, res)
    if hashlib.md5(out).digest() != hashlib.md5(ro).digest() or hashlib.md5(failed).digest() != hashlib.md5(rf).digest():
        bad += 1
        print("MISMATCH case", n, what, "out", len(out), len(ro), "failed", len(failed), len(rf))
        if bad > 5:
            break
    n += 1
print("cases compared", n, "mismatches", bad, "(refused by the reference binary:", refused, ")")
