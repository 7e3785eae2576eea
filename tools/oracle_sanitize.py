"""Development aid (CPU): the C oracle (oracle/fpl_oracle.c) built with -fsanitize=address,undefined and driven over the
whole option matrix (tests/cases.py: OPTION_SETS, MASK_BREAK_SETS x adversarial / ONT-like batches) plus N random cases —
the checker itself must be free of out-of-bounds reads and undefined behaviour before its answers pin anything.
usage: python tools/oracle_sanitize.py [n_random]        (re-executes itself with the sanitizer runtimes preloaded)"""
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/liboracle_asan.so"

if os.environ.get("FPL_ORACLE_SANITIZE_CHILD") != "1":
    subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-g", "-fPIC", "-shared", "-fsanitize=address,undefined",
                           "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"), "-o", SO,
                           os.path.join(ROOT, "oracle", "fpl_oracle.c"), "-lm"])
    libs = [subprocess.check_output(["gcc", "-print-file-name=" + n], text=True).strip() for n in ("libasan.so", "libubsan.so")]
    env = dict(os.environ, LD_PRELOAD=" ".join(libs), ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1", FPL_ORACLE_SANITIZE_CHILD="1")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
oracle_lib.build_oracle = lambda: SO
import cases  # noqa: E402
from oracle_lib import OracleEngine  # noqa: E402

n = 0
for name, opt in list(cases.OPTION_SETS.items()) + list(cases.MASK_BREAK_SETS.items()):
    for b in (cases.adversarial_batch(11), cases.ont_batch(3, n=60, mean=1500, p_chimera=0.2, p_polya=0.1)):
        o = OracleEngine(opt)
        assert o.so == SO
        o.process(b)
        if opt.mask or opt.break_reads:
            o.segments()
            o.mask_regions()
        o.stats(0, max(1, int(b.lens.max())))
        o.stats(1, max(1, int(b.lens.max())))
        o.counters()
        o.close()
        n += 1
rng = random.Random(5)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    opt, b, what = cases.random_case(rng)
    o = OracleEngine(opt)
    o.process(b)
    o.close()
    n += 1
print("oracle under ASan + UBSan: %d runs, no report" % n)
