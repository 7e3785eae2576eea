"""Output assembly on the device (device half): fpl_emit_fastq_host must hand back exactly the text
the per-pack loop of processSingleEnd appends for the two writers (src/seprocessor.cpp:264-288, Read::appendToString).
Expected text: the Python mirror of those rules (hostside.emit_fastq / emit_fastq_ext — itself held to the reference
binary's output files by tests/test_oracle_golden.py and tools/fuzz_emitter_vs_binary.py) applied to the ORACLE's
records, so a wrong record and a wrong byte both show."""
import numpy as np
import pytest

import cases
from fastplong_b200 import hostside, pack_reads
from oracle_lib import OracleEngine

pytestmark = pytest.mark.gpu


def fastq_of(batch):
    """FASTQ text with names of varying length and '+' lines that sometimes repeat the name."""
    names, plus, parts = [], [], []
    for i in range(batch.n_reads):
        s, q = batch.read(i)
        nm = b"@read%d %s" % (i, b"x" * (i % 41))
        pl = b"+" if i % 3 else b"+read%d" % i
        names.append(nm); plus.append(pl)
        parts.append(nm + b"\n" + s + b"\n" + pl + b"\n" + q + b"\n")
    return b"".join(parts), names, plus


def check(opt, batch, what, want_failed=True):
    from fastplong_b200.binding import Engine
    text, names, plus = fastq_of(batch)
    g, o = Engine(opt), OracleEngine(opt)
    got = g.process_fastq(text)
    assert got is not None, what
    recs, res, used = got
    assert used == len(text) and len(recs) == batch.n_reads
    ores = o.process(batch)
    if opt.mask or opt.break_reads:
        exp_out, exp_failed = hostside.emit_fastq_ext(batch, names, ores, o.segments(), o.mask_regions(), strand=plus)
    else:
        exp_out, exp_failed = hostside.emit_fastq(batch, names, ores, strand=plus)
    out, failed = g.emit_fastq(want_failed)
    assert len(out) == len(exp_out), f"{what}: --out has {len(out)} bytes, expected {len(exp_out)}"
    assert out == exp_out, f"{what}: --out text differs"
    if want_failed:
        assert failed == exp_failed, f"{what}: --failed_out text differs"
    else:
        assert failed == b""
    return len(exp_out), len(exp_failed)


@pytest.mark.parametrize("name", ["default_se", "cut_polyx_cplx", "trims_limits", "strict_ed0"])
def test_emit_plain_modes(name):
    n_out, n_failed = check(cases.OPTION_SETS[name], cases.ont_batch(91, n=400, mean=2500, p_chimera=0.2), name)
    assert n_out + n_failed > 0


def test_emit_every_option_set_on_the_adversarial_batch():
    b = cases.adversarial_batch(3)
    for name in sorted(cases.OPTION_SETS):
        check(cases.OPTION_SETS[name], b, f"{name}/adv")


def test_emit_without_failed_writer_and_repeat_call():
    from fastplong_b200.binding import Engine
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    b = cases.ont_batch(92, n=200, mean=1500, p_chimera=0.1)
    check(opt, b, "no failed writer", want_failed=False)
    text, names, plus = fastq_of(b)
    g = Engine(opt)
    g.process_fastq(text)
    a1 = g.emit_fastq(True)
    a2 = g.emit_fastq(True)            # the text stays built: the same bytes again
    assert a1 == a2
    g.process(pack_reads([b.read(0)]))  # any other call invalidates the chunk
    with pytest.raises(Exception):
        g.emit_fastq(True)


@pytest.mark.parametrize("name", sorted(cases.MASK_BREAK_SETS))
def test_emit_mask_break(name):
    check(cases.MASK_BREAK_SETS[name], cases.blocky_quality_batch(77), f"{name}/blocky")
    check(cases.MASK_BREAK_SETS[name], cases.adversarial_batch(6), f"{name}/adv")


def test_emit_empty_and_tiny_reads():
    reads = [(b"", b""), (b"A", b"I"), (b"ACGT" * 5, b"I" * 20), (b"N" * 40, b"#" * 40), (b"", b"")]
    check(cases.OPTION_SETS["default_se"], pack_reads(reads), "tiny")


def test_emit_long_reads_unaligned_windows():
    rng = np.random.default_rng(5)
    reads = []
    for i in range(40):
        n = int(rng.integers(30000, 90000)) + i
        reads.append((rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes(),
                      rng.integers(35, 75, n, dtype=np.uint8).tobytes()))
    check(cases.OPTION_SETS["default_se"], pack_reads(reads), "long")
