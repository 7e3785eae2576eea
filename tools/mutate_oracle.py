"""Development aid (CPU): mutation testing of the differential harness.  Every single-token mutant of oracle/fpl_oracle.c
(comparison operators, && / ||, +-1, ++) is compiled and held to the unmodified reference operators (oracle/_ref/libfplref.so)
on the test battery — the fixed option matrix x adversarial / ONT-like / blocky-quality / RNA / long-adapter / 64-entry-FASTA
batches, cases.edge_cases(), 60 random cases.  A mutant the battery cannot tell from the reference is a behaviour the tests do
not pin (or an equivalent mutant: min/max written with < or <=, unreachable guards, reads of the byte behind a read).
The survivors of the first run are what cases.RNA_SETS and cases.edge_cases() were written for.
usage: python tools/mutate_oracle.py [workers]          (about 10 minutes on 8 cores)
       python tools/mutate_oracle.py --battery <oracle.so>   (internal: one mutant)"""
import os
import random
import re
import signal
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
LINES = open(os.path.join(ROOT, "oracle", "fpl_oracle.c")).read().split("\n")
OPS = [(r"<=", "<"), (r">=", ">"), (r"(?<![<>=!])<(?![<=])", "<="), (r"(?<![<>=!-])>(?![>=])", ">="), (r"==", "!="), (r"!=", "=="),
       (r"&&", "||"), (r"\|\|", "&&"), (r"\+ 1\b", "+ 0"), (r"- 1\b", "- 0"), (r"\+\+", "--")]
TMP = "/tmp/fpl_mutants"


def battery(so):
    import oracle_lib
    oracle_lib.build_oracle = lambda: so
    import cases
    from oracle_lib import OracleEngine, RefEngine, compare_lists, compare_results, compare_stats
    signal.alarm(300)

    def check(opt, batch):
        o, r = OracleEngine(opt), RefEngine(opt)
        try:
            compare_results(o.process(batch), r.process(batch))
            if opt.mask or opt.break_reads:
                compare_lists(o.segments(), r.segments(), "segments")
                compare_lists(o.mask_regions(), r.mask_regions(), "regions")
            cyc = max(1, int(batch.lens.max()) if batch.n_reads else 1)
            for w in (0, 1):
                compare_stats(o.stats(w, cyc), r.stats(w, cyc))
            compare_stats(o.counters(), r.counters())
        finally:
            r.close()
    try:
        adv = cases.adversarial_batch(1)
        ont = cases.ont_batch(77, n=80, mean=2000, p_chimera=0.05, p_polya=0.05)
        for name, opt in cases.OPTION_SETS.items():
            if name.startswith("long_adapter_"):
                n = int(name.split("_")[-1])
                if n in (33, 129, 641):
                    check(opt, cases.long_adapter_batch(n, 900 + n, n=40))
                continue
            check(opt, adv)
            check(opt, ont)
        for name, opt in cases.MASK_BREAK_SETS.items():
            check(opt, cases.blocky_quality_batch(5, n=60))
            check(opt, adv)
        for name, opt in cases.RNA_SETS.items():
            check(opt, cases.rna_batch(41, n=80, mixed=True))
        check(cases.OPTION_SETS["fasta64_polyx"], cases.hifi_fasta64_batch(3, n=60))
        for name, (opt, b) in cases.edge_cases().items():
            check(opt, b)
        rng = random.Random(1)
        for i in range(60):
            opt, b, w = cases.random_case(rng)
            check(opt, b)
        print("SURVIVED")
    except AssertionError:
        print("KILLED")
    except Exception as e:        # the mutant made the oracle misbehave in another way (bad sizes, alarm)
        print("KILLED (%s)" % type(e).__name__)


def mutants():
    out = []
    in_comment = False
    for li, line in enumerate(LINES):
        code = line
        if in_comment:
            if "*/" not in code:
                continue
            code = " " * (code.index("*/") + 2) + code[code.index("*/") + 2:]
            in_comment = False
        if "/*" in code and "*/" not in code[code.index("/*"):]:
            in_comment = True
            code = code[:code.index("/*")]
        code = re.sub(r"/\*.*?\*/", lambda m: " " * len(m.group(0)), code).split("//")[0]
        if code.strip().startswith("#") or not code.strip():
            continue
        for pat, rep in OPS:
            for m in re.finditer(pat, code):
                out.append((li, line[:m.start()] + rep + line[m.end():]))
    return out


def run_one(args):
    k, (li, new) = args
    lines = list(LINES)
    lines[li] = new
    c, so = f"{TMP}/m{k}.c", f"{TMP}/m{k}.so"
    open(c, "w").write("\n".join(lines))
    r = subprocess.run(["gcc", "-std=gnu11", "-O1", "-fPIC", "-shared", "-w", "-I", os.path.join(ROOT, "include"), "-I",
                        os.path.join(ROOT, "oracle"), "-o", so, c, "-lm"], capture_output=True)
    if r.returncode != 0:
        return k, "nocompile"
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--battery", so], capture_output=True, text=True, timeout=600)
        out = p.stdout.strip().split("\n")[-1] if p.stdout.strip() else "KILLED (crash)"
    except subprocess.TimeoutExpired:
        out = "KILLED (timeout)"
    for f in (c, so):
        os.remove(f)
    return k, out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--battery":
        battery(sys.argv[2])
        sys.exit(0)
    os.makedirs(TMP, exist_ok=True)
    ms = mutants()
    print("mutants:", len(ms), flush=True)
    tally = {}
    with ProcessPoolExecutor(int(sys.argv[1]) if len(sys.argv) > 1 else 8) as ex:
        for k, out in ex.map(run_one, list(enumerate(ms))):
            tally[out.split(" ")[0]] = tally.get(out.split(" ")[0], 0) + 1
            if out.startswith("SURVIVED"):
                li, new = ms[k]
                print(f"SURVIVED line {li + 1}: {LINES[li].strip()[:120]}   ==>   {new.strip()[:120]}", flush=True)
    print(tally)
