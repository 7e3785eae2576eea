"""Host-side half of processSingleEnd, driven by the per-read records the C ABI returns.

This mirrors what the C++ drop-in (fastplong_b200/host/seprocessor_gpu.cpp) does with the records:
output FASTQ assembly (src/seprocessor.cpp:272-280, Read::appendToString src/read.cpp:119-173,
Read::breakByGap names src/read.cpp:192-215) and the adapter-count map
(FilterResult::addAdapterTrimmed, src/filterresult.cpp:69-77).  Pure bookkeeping: no per-base decisions.
"""
from collections import OrderedDict

import numpy as np

from . import abi


def _plus_line(strand, i):
    return strand[i] if isinstance(strand, (list, tuple)) else strand


def emit_fastq(batch, names, results, strand=b"+"):
    """Returns (out_bytes, failed_bytes) exactly as the writer threads receive them (pack order == read order).
    `strand`: the '+' line of every read, or a list with one per read (it is kept verbatim, src/read.cpp:119-143)."""
    out, failed = [], []
    seq, qual = batch.seq, batch.qual
    plus = strand
    for i in range(batch.n_reads):
        strand = _plus_line(plus, i)
        r = results[i]
        o = int(batch.offsets[i])
        nseg = int(r["n_segments"])
        name = names[i]
        for k in range(nseg):
            lo, ln = int(r["seg_lo"][k]), int(r["seg_len"][k])
            code = int(r["seg_result"][k])
            if code == abi.PASS_FILTER:
                nm = name
                if r["flags"] & abi.FLAG_MIDDLE_ADAPTER:
                    right = (k == 1) or bool(r["flags"] & abi.FLAG_SEG0_IS_RIGHT)
                    tag = b"split-by-adapter-right-" if right else b"split-by-adapter-left-"
                    nm = name[:1] + tag + name[1:]
                out.append(nm + b"\n" + seq[o + lo:o + lo + ln].tobytes() + b"\n" + strand + b"\n" +
                           qual[o + lo:o + lo + ln].tobytes() + b"\n")
            elif nseg == 1:
                tl, tn = int(r["trim_lo"]), int(r["trim_len"])
                failed.append(name + b" " + abi.FAILED_TYPES[code].encode() + b"\n" +
                              seq[o + tl:o + tl + tn].tobytes() + b"\n" + strand + b"\n" +
                              qual[o + tl:o + tl + tn].tobytes() + b"\n")
    return b"".join(out), b"".join(failed)


def emit_fastq_ext(batch, names, results, segments, regions, strand=b"+"):
    """--mask/--break variant (src/seprocessor.cpp:235-288): `segments` is the full list of output reads (read order),
    `regions` the masked regions.  Names get "r<k>-" (Read::breakByRegions) in front of the split tag."""
    out, failed = [], []
    reg_by_read = {}
    for rg in regions:
        reg_by_read.setdefault(int(rg["read"]), []).append((int(rg["lo"]), int(rg["len"])))
    ptr = 0
    plus = strand
    for i in range(batch.n_reads):
        strand = _plus_line(plus, i)
        r = results[i]
        o, L = int(batch.offsets[i]), int(batch.lens[i])
        n = int(r["n_segments"])
        pieces = segments[ptr:ptr + n]
        ptr += n
        assert all(int(p["read"]) == i for p in pieces)
        seq = batch.seq[o:o + L]
        mseq = seq
        if i in reg_by_read:
            mseq = seq.copy()
            for lo, ln in reg_by_read[i]:
                mseq[lo:lo + ln] = ord("N")
        qual = batch.qual[o:o + L]
        name = names[i]
        for p in pieces:
            lo, ln, code = int(p["lo"]), int(p["len"]), int(p["result"])
            if code == abi.PASS_FILTER:
                nm = name[1:]
                if p["split_side"]:
                    nm = (b"split-by-adapter-right-" if p["split_side"] == 2 else b"split-by-adapter-left-") + nm
                if p["break_index"]:
                    nm = b"r%d-" % int(p["break_index"]) + nm
                out.append(name[:1] + nm + b"\n" + mseq[lo:lo + ln].tobytes() + b"\n" + strand + b"\n" +
                           qual[lo:lo + ln].tobytes() + b"\n")
            elif n == 1:
                tl, tn = int(r["trim_lo"]), int(r["trim_len"])
                src = mseq if p["is_r1"] else seq
                failed.append(name + b" " + abi.FAILED_TYPES[code].encode() + b"\n" + src[tl:tl + tn].tobytes() + b"\n" +
                              strand + b"\n" + qual[tl:tl + tn].tobytes() + b"\n")
    assert ptr == len(segments)
    return b"".join(out), b"".join(failed)


def default_names(batch, prefix=b"read"):
    return [b"@%s%d len=%d" % (prefix, i, int(batch.lens[i])) for i in range(batch.n_reads)]


def adapter_count_map(counters, adapters):
    """FilterResult::mAdapter from the device event table, ordered like classcomp (length, then lexicographic;
    src/filterresult.h:14-23)."""
    m = {}
    width = abi.MAX_ADAPTER_LEN + 1
    table = np.asarray(counters[abi.CNT_FIXED:abi.CNT_FIXED + len(adapters) * 2 * width]).reshape(len(adapters), 2, width)
    for idx, side, c in zip(*np.nonzero(table)):
        a = adapters[idx]
        s = a[len(a) - c:] if side == 0 else a[:c]
        if not s:
            continue
        m[s] = m.get(s, 0) + int(table[idx, side, c])
    return OrderedDict(sorted(m.items(), key=lambda kv: (len(kv[0]), kv[0])))


def report_summary(pre, post, counters, cycles):
    """The scalar JSON fields the reference derives from Stats/FilterResult (Stats::summarize src/stats.cpp:150-256,
    FilterResult::reportJson src/filterresult.cpp:120-132), for comparison with golden JSON reports."""
    def summ(block):
        C_ = cycles
        content = block[:8 * C_].reshape(8, C_)
        tail = block[16 * C_:]
        total_per_cycle = content.sum(axis=0)
        nz = np.nonzero(total_per_cycle == 0)[0]
        ncyc = int(nz[0]) if len(nz) else C_          # :155-163 stops at the first zero cycle
        bases = int(total_per_cycle[:ncyc].sum())
        qh = tail[abi.STATS_QUALHIST:abi.STATS_QUALHIST + 128]
        reads = int(tail[abi.STATS_READS])
        return {"total_reads": reads, "total_bases": bases, "q20_bases": int(qh[33 + 20:127].sum()),
                "q30_bases": int(qh[33 + 30:127].sum()), "total_cycles": ncyc,
                "read_mean_length": int(tail[abi.STATS_LENSUM]) // reads if reads else 0}
    f = counters[abi.CNT_FILTER:abi.CNT_FILTER + 32]
    return {"before": summ(pre), "after": summ(post),
            "filtering_result": {"passed_filter_reads": int(f[abi.PASS_FILTER]),
                                 "low_quality_reads": int(f[abi.FAIL_QUALITY]),
                                 "too_many_N_reads": int(f[abi.FAIL_N_BASE]),
                                 "too_short_reads": int(f[abi.FAIL_LENGTH]),
                                 "too_long_reads": int(f[abi.FAIL_TOO_LONG])},
            "adapter_trimmed_reads": int(counters[abi.CNT_ADAPTER_READS]),
            "adapter_trimmed_bases": int(counters[abi.CNT_ADAPTER_BASES])}
