"""Development aid (CPU): the kernels' own source, executed under the SIMT emulator (tests/simt_emu.py), against the oracle on
random option sets x seeded batches — the three families of tests/cases.py, as packed batches and as FASTQ text through the
device parser and the device output assembly (the text against the Python mirror of the output rules on the oracle's records).
The CPU twin of tools/fuzz_gpu_vs_oracle.py.   usage: python tools/fuzz_emulated_vs_oracle.py <seed> <seconds>"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import test_simt_kernels as T  # noqa: E402

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
t0 = time.time()
n = bad = 0
tally = {}
while time.time() - t0 < budget:
    family = rng.choice(["random_case"] * 6 + ["random_case_many_adapters"] * 2 + ["random_case_long_reads"])
    opt, batch, what = getattr(cases, family)(rng)
    mode = rng.choice(["packed/jit", "packed/fast", "packed/generic", "text"])
    try:
        if mode == "text":
            T.check_text(opt, batch, what, want_failed=rng.random() < 0.8, last_newline=rng.random() < 0.7)
        else:
            T.check(opt, batch, what, mode.split("/")[1])
    except Exception as e:      # a mismatch (AssertionError) or an emulator abort surfaced as an error
        bad += 1
        print("MISMATCH" if isinstance(e, AssertionError) else "ERROR", "case", n, family, mode, what, str(e)[:300], flush=True)
        if bad > 5:
            break
    tally[mode] = tally.get(mode, 0) + 1
    n += 1
print("cases", n, tally, "mismatches", bad)
