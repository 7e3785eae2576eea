"""Device-side FASTQ ingest: fpl_process_fastq_host must find exactly the records the reference's
FastqReader finds in a well-formed file, produce the same per-read results as the packed-batch entry point, and refuse
(return 1) anything outside the strict layout."""
import numpy as np
import pytest

import cases
from fastplong_b200 import pack_reads, synth
from oracle_lib import compare_results, compare_stats

pytestmark = pytest.mark.gpu


def engine(opt):
    from fastplong_b200.binding import Engine
    return Engine(opt)


def fastq_text(reads, plus=b"+", last_newline=True):
    parts = []
    for i, (s, q) in enumerate(reads):
        parts.append(b"@r%d some description\n%s\n%s\n%s\n" % (i, s, plus if i % 3 else b"+r%d" % i, q))
    t = b"".join(parts)
    return t if last_newline else t[:-1]


def test_ingest_matches_packed_entry_point():
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    b = cases.ont_batch(41, n=500, mean=3000, p_chimera=0.05)
    reads = [b.read(i) for i in range(b.n_reads)] + [(b"", b""), (b"ACGT", b"IIII")]
    text = fastq_text(reads)
    g1, g2 = engine(opt), engine(opt)
    out = g1.process_fastq(text)
    assert out is not None
    recs, res, used = out
    assert used == len(text) and len(recs) == len(reads)
    for i in (0, 1, len(reads) - 2, len(reads) - 1):
        r = recs[i]
        assert text[r["seq_off"]:r["seq_off"] + r["seq_len"]] == reads[i][0]
        assert text[r["qual_off"]:r["qual_off"] + r["seq_len"]] == reads[i][1]
        assert text[r["name_off"]:r["name_off"] + r["name_len"]] == b"@r%d some description" % i
        assert text[r["plus_off"]:r["plus_off"] + r["plus_len"]].startswith(b"+")
    ref = g2.process(pack_reads(reads))
    compare_results(res, ref, "ingest")
    cyc = max(g1.cycles, g2.cycles)
    g1.reserve_cycles(cyc); g2.reserve_cycles(cyc)
    for w in (0, 1):
        compare_stats(g1.stats(w), g2.stats(w), f"ingest/stats{w}")
    compare_stats(g1.counters(), g2.counters(), "ingest/counters")


def test_ingest_chunking_and_unterminated_last_line():
    opt = cases.OPTION_SETS["default_se"]
    b = cases.ont_batch(42, n=120, mean=1500)
    reads = [b.read(i) for i in range(b.n_reads)]
    text = fastq_text(reads, last_newline=False)
    whole = engine(opt)
    ref = whole.process(pack_reads(reads))
    g = engine(opt)
    got, pos, carry = [], 0, b""
    chunk = 70001
    while pos < len(text) or carry:
        piece = carry + text[pos:pos + chunk]
        pos += chunk
        last = pos >= len(text)
        out = g.process_fastq(piece, is_last=last)
        assert out is not None
        recs, res, used = out
        got.append(res)
        carry = piece[used:]
        if last:
            assert used == len(piece)
            break
    compare_results(np.concatenate(got), ref, "chunked ingest")
    compare_stats(g.counters(), whole.counters(), "chunked/counters")


@pytest.mark.parametrize("kind", ["crlf", "blank_line", "no_at", "no_plus", "len_mismatch", "three_lines"])
def test_ingest_refuses_non_strict_layouts(kind):
    reads = [(b"ACGTACGTAC", b"IIIIIIIIII"), (b"GGGTTTAAAC", b"IIIIIIIIII")]
    text = fastq_text(reads)
    if kind == "crlf":
        text = text.replace(b"\n", b"\r\n")
    elif kind == "blank_line":
        text = text.replace(b"IIIIIIIIII\n@r1", b"IIIIIIIIII\n\n@r1")
    elif kind == "no_at":
        text = text.replace(b"@r1", b"r1x")
    elif kind == "no_plus":
        text = text.replace(b"\n+\n", b"\n-\n")
    elif kind == "len_mismatch":
        text = text.replace(b"GGGTTTAAAC", b"GGGTTTAAA")
    elif kind == "three_lines":
        text = text[: text.rfind(b"\n", 0, len(text) - 1) + 1]
    g = engine(cases.OPTION_SETS["default_se"])
    assert g.process_fastq(text) is None
    assert not g.counters().any()     # nothing was accumulated
