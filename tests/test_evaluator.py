"""Adapter auto-detection: fastplong_b200/evaluator.py (host half) against the unmodified reference
binary's own detection, and the device ten-mer tables against their numpy restatement."""
import json
import os
import subprocess

import numpy as np
import pytest

from fastplong_b200 import evaluator, synth
from oracle_lib import REF_BIN, kmer10_tables


def reference_detection(batch, tmp_path, extra=()):
    fq = str(tmp_path / "in.fq")
    synth.to_fastq(batch, fq)
    js = str(tmp_path / "ref.json")
    r = subprocess.run([REF_BIN, "-i", fq, "-o", str(tmp_path / "o.fq"), "-j", js, "-h", str(tmp_path / "r.html"), "-w", "2", *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    j = json.load(open(js))
    ac = j.get("adapter_cutting", {})
    return ac.get("read_start_adapter"), ac.get("read_end_adapter"), r.stderr


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/fastplong_ref not built")
@pytest.mark.parametrize("seed,kw", [(5, {}), (6, dict(p_start=0.5, p_end=0.9)), (7, dict(p_start=0.05, p_end=0.6)),
                                     (8, dict(adapter_start="GATCGGAAGAGCACACGTCTGAACTCCAGTCAC", adapter_end="ACACTCTTTCCCTACACGACGCTCTTCCGATCT"))])
def test_host_logic_matches_reference_binary(seed, kw, tmp_path):
    batch = synth.ont_like(600, 1500, seed, **kw)
    start, end, log = reference_detection(batch, tmp_path)
    got = evaluator.detect_adapters(batch, kmers=kmer10_tables)
    # the JSON prints "unspecified" for an adapter left at "auto" (src/options.cpp:247-259)
    norm = lambda s: "unspecified" if s == "auto" else s
    assert (norm(got[0]), norm(got[1])) == (start, end), log[-600:]


def test_fewer_than_100_reads_detects_nothing():
    batch = synth.ont_like(60, 800, 3)
    assert evaluator.detect_adapters(batch, kmers=kmer10_tables) == ("auto", "auto")


def test_evaluated_prefix_limits():
    assert evaluator.evaluated_prefix([10] * 5) == 5
    assert evaluator.evaluated_prefix([1000] * 70000) == 64 * 1024
    lens = [300_000_000, 300_000_000, 5, 5]
    assert evaluator.evaluated_prefix(lens) == 2          # the read that crosses 512 Mi bases is still loaded


@pytest.mark.gpu
@pytest.mark.parametrize("side", [0, 1])
def test_device_tables_match_numpy(side):
    from fastplong_b200.binding import eval_adapter_kmers
    from fastplong_b200 import pack_reads
    reads = synth.adversarial_reads(9) + [synth.ont_like(300, 900, 4).read(i) for i in range(300)]
    batch = pack_reads(reads)
    for shift in (1, 3):
        c, a, t = eval_adapter_kmers(batch, side, shift)
        rc, ra, rt = kmer10_tables(batch, side, shift)
        assert t == rt and np.array_equal(c, rc) and np.array_equal(a, ra)


@pytest.mark.gpu
def test_device_detection_finds_the_planted_adapters():
    batch = synth.ont_like(800, 2000, 21)
    assert evaluator.detect_adapters(batch) == (synth.ADAPTER_START, synth.ADAPTER_END)
