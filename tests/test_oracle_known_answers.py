"""The oracle (oracle/fpl_oracle.c) against every known-answer vector the reference's own tests hold for the
hot path.  CPU only."""
import pytest

from fastplong_b200 import Options, pack_reads
from oracle_lib import OracleEngine, RefEngine, have_ref

ENGINES = [OracleEngine] + ([RefEngine] if have_ref() else [])


@pytest.mark.parametrize("Engine", ENGINES)
def test_trim_by_sequence_start_and_end(Engine):
    # test/adaptertrimmer_test.cpp:4-18 (edMax 0.3, trimmingExtension 0)
    adapter = "GCGCATACTTTTCCACGGGGATACTACTG"
    r = (b"AGGTGCTGCGCATACTTTTCCACGGGGATACTACTGGGTGTTACCGTGGGAATGAATCCTTTTAACCTTAGCAATACGTAAAGGTGCT",
         b"///EEEEEEEEEEEEEEEEEEEEEEEEEE////EEEEEEEEEEEEE////E////EEEEEEEEE///EEEEEEEEEEEEEEEEEEEEE")
    opt = Options(start_adapter=adapter, end_adapter="", distance_threshold=0.3, trimming_extension=0,
                  disable_quality_filtering=True, disable_length_filtering=True)
    res = Engine(opt).process(pack_reads([r]))[0]
    assert r[0][res["trim_lo"]:res["trim_lo"] + res["trim_len"]] == b"GGTGTTACCGTGGGAATGAATCCTTTTAACCTTAGCAATACGTAAAGGTGCT"
    r2 = (b"TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAAGCGCATACTTTTCCACGGGGA",
          b"///EEEEEEEEEEEEEEEEEEEEEEEEEE////EEEEEEEEEEEEE////E////EEEEEEEEET")
    opt = Options(start_adapter="", end_adapter=adapter, distance_threshold=0.3, trimming_extension=0,
                  disable_quality_filtering=True, disable_length_filtering=True)
    res = Engine(opt).process(pack_reads([r2]))[0]
    assert r2[0][res["trim_lo"]:res["trim_lo"] + res["trim_len"]] == b"TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAA"


@pytest.mark.parametrize("Engine", ENGINES)
def test_search_adapter_left(Engine):
    # test/adaptertrimmer_test.cpp:37-57
    read = b"TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTTAAAATTTTCCCCGGGGAAATTTCCCGGGAAATTTCCCGGGATCGATCGATCGATCGAATTCC"
    e = Engine(Options(distance_threshold=0.3))
    assert e.search_adapter(read, b"TTTT", 0, -1, True, False) == 0
    assert e.search_adapter(read, b"AACC", 0, -1, True, False) == 4


def test_trim_and_cut():
    # test/filter_test.cpp:4-22: cut_front + cut_tail W=4 Q20 with tail=1
    opt = Options(cut_front=True, cut_tail=True, cut_window_size=4, cut_mean_quality=20, trim_tail=1)
    seq = b"TTTTAACCCCCCCCCCCCCCCCCCCCCCCCCCCCAATTTT"
    qual = b"/////CCCCCCCCCCCC////CCCCCCCCCCCCCC////E"
    lo, n = OracleEngine(opt).trim_and_cut(seq, qual)
    assert seq[lo:lo + n] == b"CCCCCCCCCCCCCCCCCCCCCCCCCCCC"
    assert qual[lo:lo + n] == b"CCCCCCCCCCC////CCCCCCCCCCCCC"


def test_trim_polyx():
    # test/polyx_test.cpp:4-17
    seq = b"ATTTTAAAAAAAAAATAAAAAAAAAAAAACAAAAAAAAAAAAAAAAAAAAAAAAAT"
    nl, base, plen = OracleEngine(Options()).trim_polyx(seq, 10)
    assert seq[:nl] == b"ATTTT" and plen == 51 and base == 0


def test_polyx_runs_off_the_front():
    # a 30xA read is resized to length 0 with event (A, 30)
    nl, base, plen = OracleEngine(Options()).trim_polyx(b"A" * 30, 10)
    assert (nl, base, plen) == (0, 0, 30)


@pytest.mark.parametrize("Engine", ENGINES)
def test_edit_distance_vectors(Engine):
    # src/editdistance.cpp:141-172
    s1 = [b"CCTATCAGGGAGCTGTGGGCCAGCCAGGAGGCAGCACATGCCCAATCCCAGGCCCCTCCCGTTGTAAGTTCCCGTTCTACCCGACAGGGACCTGCTGACAAAAGACAGGGCTGGAGAGCCAGCCTGAAGGCCCTGGGACCCTTCTATCCAC",
          b"ACTTATGTTTTTAAATGAGGATTATTGATAGTACTCTTGGTTTTTATACCATTCAGATCACTGAATTTATAAAGTACCCATCTAGTACTTCAAAAAGTAAAGTGTTCTGCCAGATCTTAGGTATAGAGGACCCTAACACAGTAAGATCGGA",
          b"TAGGGGTATGAGTAGAGCTGAGCTGGGGGAAAAGAGGGAAATTCCCAGGGGTGGAGGAAGAGTCAAGTCCCCCTCTACACCTAGAGGATGAACTTAAGGAAGGAGTGAAGGTCATATGTGTTGTTCCTGAGGAAAAGGCCGCTGTAGAAAA"]
    s2 = [s1[0],
          b"ACTTATGTTTTTAAATGAGGATTATTGATAGTACTCTTGGTTTTTATACCATTCAGATCACTGAATTTATAAAGTACCCATCTAGTACTTGAAAAAGTAAAGTGTTCTGCCAGATCTTAGGTATAGAGGACCCTAACACAGTAAGATCGGA",
          b"CCTGGGCCTGGCCCTTGTCTAAAACTGACTCTTTTGAGGGTGATTTTGGATGTTCTTAGTAGAGTCTCTCACCTGTACTTTCCTTGCCTAAGGTGCTGTCTTCTCTTGCAGGTTGCCTACACGTTCCTCACATGCCCTAAGAACCATGGGA"]
    e = Engine(Options())
    assert [e.edit_distance(a, b) for a, b in zip(s1, s2)] == [0, 1, 90]
