"""Per-basic-block shares of executed instructions and stall samples from an ncu --import-source capture.
usage: python tools/ncu_blocks.py report.ncu-rep <kernel ordinal, 1-based> [min share]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
r = [x for x in csv.reader(out.splitlines())]
h = r[1]; ie = h.index('Instructions Executed'); src = h.index('Source'); smp = h.index('# Samples')
rows = [x for x in r[2:] if len(x) > max(ie, smp) and x[0].startswith('0x')]
tot = sum(int(x[ie]) for x in rows); ts = sum(int(x[smp]) for x in rows)
print(r[0][1][:70], len(rows), "SASS instr;", tot, "warp-instr executed")
i = 0
while i < len(rows):
    c = int(rows[i][ie]); j = i; e = 0; s = 0
    while j < len(rows) and abs(int(rows[j][ie]) - c) <= 0.03 * max(c, 1):
        e += int(rows[j][ie]); s += int(rows[j][smp]); j += 1
    ops = {}
    for x in rows[i:j]:
        t = x[src].split()
        o = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[o] = ops.get(o, 0) + 1
    if e / tot > thr or s / ts > thr:
        print(f"{i:4d}-{j-1:4d} n={j-i:3d} exec/instr={c:>10d} share={e/tot:6.3f} samples={s/ts:6.3f}", dict(sorted(ops.items(), key=lambda kv: -kv[1])[:7]))
    i = j
