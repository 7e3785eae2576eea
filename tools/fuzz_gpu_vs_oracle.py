"""Development aid (run under gpurun): random option sets x seeded batches (tests/cases.py:random_case), GPU library vs
the oracle — every record, both Stats blocks, the counters and the --mask/--break lists.
usage: python tools/fuzz_gpu_vs_oracle.py <seed> <seconds> [mixed|many|long]
  mixed (default): cases.random_case; many: 6-40 FASTA adapters (k_trim's pre-filter); long: reads of 60-400 kb"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle_lib import OracleEngine, compare_lists, compare_results, compare_stats  # noqa: E402
from fastplong_b200.binding import Engine

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
FAMILY = {"mixed": cases.random_case, "many": cases.random_case_many_adapters,
          "long": cases.random_case_long_reads}[sys.argv[3] if len(sys.argv) > 3 else "mixed"]
t0 = time.time()
n = bad = 0
while time.time() - t0 < budget:
    opt, batch, what = FAMILY(rng)
    r = None
    try:
        o, r = OracleEngine(opt), Engine(opt)
        compare_results(o.process(batch), r.process(batch), "records")
        if opt.mask or opt.break_reads:
            compare_lists(o.segments(), r.segments(), "segments")
            compare_lists(o.mask_regions(), r.mask_regions(), "regions")
        cyc = max(1, int(batch.lens.max()))
        for w in (0, 1):
            compare_stats(o.stats(w, cyc), r.stats(w, cyc), f"stats{w}")
        compare_stats(o.counters(), r.counters(), "counters")
    except Exception as e:     # a mismatch (AssertionError) or a library error
        bad += 1
        print("MISMATCH" if isinstance(e, AssertionError) else "ERROR", "case", n, what, type(e).__name__, str(e)[:300])
        if bad > 5:
            break
    finally:
        if r is not None and hasattr(r, "close"):
            r.close()
    n += 1
print("cases", n, "mismatches", bad)
