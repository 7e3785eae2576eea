"""tests/stats_tables.py (the torch restatement of the Stats block that bench.py's full-scale check uses on the GPU) pinned
to the C oracle on the CPU: every word of the pre- and post-filter blocks and every per-read / per-segment median."""
import numpy as np
import pytest
import torch

import cases
from fastplong_b200 import Options, abi, pack_reads, synth
from oracle_lib import OracleEngine
from stats_tables import passing_segments, stats_block


def _check(opt, batch, chunk):
    o = OracleEngine(opt)
    res = o.process(batch)
    cap = max(8, int(batch.lens.max()))
    seq, qual = torch.from_numpy(batch.seq), torch.from_numpy(batch.qual)
    pre, med = stats_block(torch, seq, qual, batch.offsets, batch.lens, cap, chunk=chunk)
    exp = o.stats(0, cap)
    assert np.array_equal(pre, exp), np.nonzero(pre != exp)[0][:8]
    nz = batch.lens > 0
    assert np.array_equal(med[nz], res["pre_median_qual"][nz])
    rd, k, starts, lens = passing_segments(res, batch.offsets)
    post, med = stats_block(torch, seq, qual, starts, lens, cap, chunk=chunk)
    exp = o.stats(1, cap)
    assert np.array_equal(post, exp), np.nonzero(post != exp)[0][:8]
    nz = lens > 0
    assert np.array_equal(med[nz], res["seg_median_qual"][rd, k][nz])
    o.close()
    return int(len(rd))


@pytest.mark.parametrize("name,chunk", [("cut_polyx_cplx", 1 << 26), ("default_se", 1 << 26), ("trims_limits", 1 << 26),
                                        ("no_adapter_no_filters", 1 << 26), ("fasta5", 1 << 26),
                                        ("cut_polyx_cplx", 4099), ("no_adapter_no_filters", 16411)])
def test_block_equals_oracle_on_adversarial_reads(name, chunk):
    """Adversarial reads: N runs, lower-case and non-ACGT letters, reads of 0..5 bases, split reads; a chunk size that is
    not a multiple of anything puts chunk borders inside reads, 5-mers and slot padding."""
    assert _check(cases.OPTION_SETS[name], cases.adversarial_batch(31), chunk) > 0


def test_block_equals_oracle_on_ont_like_reads_with_chimeras():
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    assert _check(opt, cases.ont_batch(5, n=400, mean=3000, p_chimera=0.2, p_polya=0.05), 400009) > 300


def test_block_of_tiny_and_empty_segments():
    reads = [(b"ACGTACGTACGTACGTACGT"[:n], bytes([33 + 30]) * n) for n in (0, 1, 2, 3, 4, 5, 6, 9)] * 3
    _check(Options(disable_adapter_trimming=True, length_required=0), pack_reads(reads), 7)
    b = pack_reads([(b"", b"")] * 3)
    blk, med = stats_block(torch, torch.from_numpy(b.seq), torch.from_numpy(b.qual), b.offsets, b.lens, 8)
    assert blk[16 * 8 + abi.STATS_READS] == 3 and blk.sum() == 3 and not med.any()


def test_block_rejects_overlapping_segments():
    b = synth.ont_like(4, 300, 3)
    with pytest.raises(AssertionError):
        stats_block(torch, torch.from_numpy(b.seq), torch.from_numpy(b.qual), np.array([0, 10]), np.array([50, 50]), 64)


# ---- bench.py's full_scale_check, driven on the CPU by the oracle standing in for the library ----
class _OracleBehindEngineCalls:
    """The four calls full_scale_check makes on binding.Engine, answered by the oracle."""

    def __init__(self, opt, batch, cap):
        o = OracleEngine(opt)
        self.res = o.process(batch)
        self.cycles = cap
        self.blocks = [o.stats(0, cap), o.stats(1, cap)]
        self.cnt = o.counters()
        o.close()

    def fetch_results(self, n):
        return self.res[:n].copy()

    def counters(self):
        return self.cnt.copy()

    def stats(self, which):
        return self.blocks[which].copy()


def test_bench_full_scale_check_passes_and_catches_corruption(monkeypatch):
    import bench
    from fastplong_b200 import synth_fast
    monkeypatch.setattr(bench, "FULL_CHECK_BASES", 400_000)
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    tile = synth_fast.ont_like_device(600, 2000, 77, "cpu")
    batch = tile.to_host()
    eng = _OracleBehindEngineCalls(opt, batch, 1 << int(np.ceil(np.log2(batch.lens.max()))))
    args = (torch, eng, opt, tile, tile.seq, tile.qual, tile.offsets, tile.lens, tile.n_reads, 2000)
    out = bench.full_scale_check(*args)
    assert out["ok"] is True and out["records_vs_oracle"]["read_ranges"][-1][1] == tile.n_reads
    assert out["records_vs_oracle"]["reads"] >= 8 * 4 and out["stats_vs_torch"]["passing_segments"] > 400
    # a wrong record in the last range, a wrong Stats word, a wrong median, a wrong counter: each one is caught
    good = eng.res["trim_len"][-1]
    eng.res["trim_len"][-1] += 1
    with pytest.raises(AssertionError, match="full-scale records"):
        bench.full_scale_check(*args)
    eng.res["trim_len"][-1] = good
    eng.blocks[1][12345 % len(eng.blocks[1])] += 1
    with pytest.raises(AssertionError, match="Stats block 1"):
        bench.full_scale_check(*args)
    eng.blocks[1][12345 % len(eng.blocks[1])] -= 1
    eng.cnt[abi.CNT_FILTER + abi.FAIL_LENGTH] += 1
    with pytest.raises(AssertionError, match="filter results"):
        bench.full_scale_check(*args)
    eng.cnt[abi.CNT_FILTER + abi.FAIL_LENGTH] -= 1
    with pytest.raises(TimeoutError):
        bench.full_scale_check(*args, budget_s=-1.0)
    assert bench.full_scale_check(*args)["ok"] is True
