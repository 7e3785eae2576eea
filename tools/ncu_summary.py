#!/usr/bin/env python
"""Summarise an .ncu-rep (development aid): per-kernel headline metrics + opcode mix from the source page.
usage: python tools/ncu_summary.py report.ncu-rep [--ops]"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "?")[:70], "id", d.get("ID"))
        for w in WANT:
            if w in d:
                print("   %-70s %s %s" % (w, d[w], rows[1][hdr.index(w)]))
        stalls = [(float(d[k].replace(",", "") or 0), k) for k in hdr
                  if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k]]
        for v, k in sorted(stalls, reverse=True)[:6]:
            print("   stall %-55s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


def ops(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, None
    def flush():
        if cur and agg:
            tot = sum(agg.values())
            print("== ops", cur[:70], "total warp-instr", tot)
            print("   " + "  ".join(f"{k}:{v / tot:.3f}" for k, v in agg.most_common(16)))
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            flush()
            cur, agg, hdr = r[1], collections.Counter(), None
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            try:
                n = int(r[hdr.index("Instructions Executed")])
            except ValueError:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[hdr.index("Source")])
            agg[(m.group(2).split(".")[0] if m else "?")] += n
    flush()


if __name__ == "__main__":
    raw(sys.argv[1])
    if "--ops" in sys.argv:
        ops(sys.argv[1])
