"""Development aid (run under gpurun): random option sets x seeded batches (tests/cases.py:random_case), GPU library vs
the oracle — every record, both Stats blocks, the counters and the --mask/--break lists.
usage: python tools/fuzz_gpu_vs_oracle.py <seed> <seconds> [mixed|many|long|text]
  mixed (default): cases.random_case; many: 6-40 FASTA adapters (k_trim's pre-filter); long: reads of 60-400 kb;
  text: the mixed cases as FASTQ text through fpl_process_fastq_host + fpl_emit_fastq_host (records and both output texts)"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle_lib import OracleEngine, compare_lists, compare_results, compare_stats  # noqa: E402
from fastplong_b200 import hostside  # noqa: E402
from fastplong_b200.binding import Engine

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
MODE = sys.argv[3] if len(sys.argv) > 3 else "mixed"
FAMILY = {"mixed": cases.random_case, "many": cases.random_case_many_adapters,
          "long": cases.random_case_long_reads, "text": cases.random_case}[MODE]


def fastq_of(batch, rng):
    names, plus, parts = [], [], []
    for i in range(batch.n_reads):
        sq, q = batch.read(i)
        nm = b"@r%d %s" % (i, b"d" * rng.randrange(0, 40))
        pl = b"+" if rng.random() < 0.7 else b"+" + nm[1:]
        names.append(nm); plus.append(pl)
        parts.append(nm + b"\n" + sq + b"\n" + pl + b"\n" + q + b"\n")
    return b"".join(parts), names, plus


t0 = time.time()
n = bad = 0
while time.time() - t0 < budget:
    opt, batch, what = FAMILY(rng)
    r = None
    try:
        o, r = OracleEngine(opt), Engine(opt)
        if MODE == "text":
            text, names, plus = fastq_of(batch, rng)
            ores = o.process(batch)
            got = r.process_fastq(text)
            assert got is not None and got[2] == len(text), "strict FASTQ refused"
            compare_results(ores, got[1], "records")
            if opt.mask or opt.break_reads:
                exp = hostside.emit_fastq_ext(batch, names, ores, o.segments(), o.mask_regions(), strand=plus)
            else:
                exp = hostside.emit_fastq(batch, names, ores, strand=plus)
            out = r.emit_fastq(True)
            assert out[0] == exp[0], "--out text differs"
            assert out[1] == exp[1], "--failed_out text differs"
        else:
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
