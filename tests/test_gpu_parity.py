"""Parity tests proper: the CUDA path (through the C ABI, libfplgpu.so) against the oracle and the committed
golden fixtures, bit-exact on every record field, both Stats blocks and the FilterResult counters."""
import os

import numpy as np
import pytest

import cases
from fastplong_b200 import Options, PackedBatch, pack_reads, synth
from oracle_lib import OracleEngine, compare_results, compare_stats

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_engine(opt):
    from fastplong_b200.binding import Engine
    return Engine(opt)


def check_against_oracle(opt, batch, what):
    g, o = gpu_engine(opt), OracleEngine(opt)
    compare_results(g.process(batch), o.process(batch), what)
    cyc = max(1, int(batch.lens.max()) if batch.n_reads else 1)
    for w in (0, 1):
        compare_stats(g.stats(w, cyc), o.stats(w, cyc), f"{what}/stats{w}")
    compare_stats(g.counters(), o.counters(), what + "/counters")
    g.close()


@pytest.mark.parametrize("name", sorted(cases.OPTION_SETS))
def test_golden_fixture(name):
    z = np.load(os.path.join(GOLDEN, "adversarial_input.npz"))
    b = PackedBatch(z["seq"], z["qual"], z["offsets"], z["lens"])
    ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    g = gpu_engine(cases.OPTION_SETS[name])
    compare_results(g.process(b), ref["results"], name)
    cyc = int(ref["cycles"])
    compare_stats(g.stats(0, cyc), ref["pre"], name + "/pre")
    compare_stats(g.stats(1, cyc), ref["post"], name + "/post")
    compare_stats(g.counters(), ref["counters"], name + "/counters")


@pytest.mark.parametrize("name", sorted(cases.OPTION_SETS))
@pytest.mark.parametrize("seed", [3, 4])
def test_adversarial_vs_oracle(name, seed):
    check_against_oracle(cases.OPTION_SETS[name], cases.adversarial_batch(seed), f"{name}/adv{seed}")


@pytest.mark.parametrize("name", sorted(cases.OPTION_SETS))
def test_ont_like_vs_oracle(name):
    check_against_oracle(cases.OPTION_SETS[name],
                         cases.ont_batch(91, n=400, mean=3000, p_chimera=0.05, p_polya=0.05), name + "/ont")


@pytest.mark.parametrize("seed", [5, 6])
def test_config3_shape_fasta64_vs_oracle(seed):
    """BASELINE configs[2]: HiFi-like reads, 64-entry adapter FASTA (trimByMultiSequences over all 64 in map order) +
    polyX trim, entries planted at the read starts."""
    check_against_oracle(cases.OPTION_SETS["fasta64_polyx"], cases.hifi_fasta64_batch(seed), f"fasta64/hifi{seed}")


@pytest.mark.parametrize("n", sorted(cases.LONG_ADAPTERS))
def test_long_adapters_vs_oracle(n):
    """-s/-e adapters of 31..128 bp: every halo-word / counter-plane class of k_scan_jit and k_scan_fast."""
    check_against_oracle(cases.OPTION_SETS[f"long_adapter_{n}"], cases.long_adapter_batch(n, 900 + n), f"long{n}")


@pytest.mark.parametrize("n", [30, 64, 128])
def test_precompiled_scan_matches_jit(n, monkeypatch):
    """FPL_NO_JIT=1 selects the precompiled k_scan_fast; FPL_FORCE_GENERIC_SCAN=1 the byte-wise k_scan."""
    if n == 30:
        opt, batch = cases.OPTION_SETS["default_se"], cases.ont_batch(17, n=300, mean=2500, p_chimera=0.1)
    else:
        opt, batch = cases.OPTION_SETS[f"long_adapter_{n}"], cases.long_adapter_batch(n, 5)
    ref = gpu_engine(opt).process(batch)
    monkeypatch.setenv("FPL_NO_JIT", "1")
    compare_results(gpu_engine(opt).process(batch), ref, "k_scan_fast")
    monkeypatch.setenv("FPL_FORCE_GENERIC_SCAN", "1")
    compare_results(gpu_engine(opt).process(batch), ref, "k_scan")


def test_config1_shape_vs_oracle():
    """BASELINE config 1 shape (ONT reads, mean 8 kb, known 30 bp adapters, default filters), 1500 reads."""
    opt = Options(start_adapter=synth.ADAPTER_START)
    check_against_oracle(opt, synth.ont_like(1500, 8000, 2024), "c1")


def test_empty_and_tiny_batches():
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    check_against_oracle(opt, pack_reads([]), "empty")
    check_against_oracle(opt, pack_reads([(b"", b"")]), "one-empty-read")
    check_against_oracle(opt, pack_reads([(b"ACGT" * 5, b"IIII" * 5)]), "one-read")


def test_long_reads_cross_tiles():
    """Reads longer than the scan tile (4096) and the stats tile (1024 cycles), incl. a 300 kb read."""
    rng = np.random.default_rng(8)
    reads = []
    for L in (4095, 4096, 4097, 8192, 12289, 70000, 300000):
        s = bytearray(synth.BASES[rng.integers(0, 4, size=L)].tobytes())
        a = synth.ADAPTER_START.encode()
        p = L // 2
        s[p:p + len(a)] = a
        q = (np.rint(rng.normal(20, 8, size=L)).clip(1, 50).astype(np.uint8) + 33).tobytes()
        reads.append((bytes(s), q))
    check_against_oracle(Options(start_adapter=synth.ADAPTER_START, low_complexity_filter=True), pack_reads(reads), "long")


def test_accumulates_across_batches_and_tiles(monkeypatch):
    """Two submissions + forced multi-tile processing must equal one oracle pass over the concatenation."""
    monkeypatch.setenv("FPL_TILE_MBASES", "1")
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    b1, b2 = cases.ont_batch(5, n=700, mean=3000), cases.ont_batch(6, n=300, mean=9000)
    g, o = gpu_engine(opt), OracleEngine(opt)
    r1, r2 = g.process(b1), g.process(b2)
    compare_results(r1, o.process(b1), "b1")
    compare_results(r2, o.process(b2), "b2")
    cyc = int(max(b1.lens.max(), b2.lens.max()))
    for w in (0, 1):
        compare_stats(g.stats(w, cyc), o.stats(w, cyc), f"acc/stats{w}")
    compare_stats(g.counters(), o.counters(), "acc/counters")
    g.reset()
    assert not g.counters().any() and not g.stats(0).any()


def test_size_independent_properties_large():
    """At a size the oracle would take minutes for: invariants that must hold for any input."""
    opt = Options(start_adapter=synth.ADAPTER_START, cut_tail=True)
    b = synth.ont_like(20000, 8000, 77)
    g = gpu_engine(opt)
    res = g.process(b)
    cyc = int(b.lens.max())
    pre, post, cnt = g.stats(0, cyc), g.stats(1, cyc), g.counters()
    from fastplong_b200 import abi
    assert pre[16 * cyc + abi.STATS_READS] == b.n_reads
    assert pre[16 * cyc + abi.STATS_LENSUM] == b.n_bases
    assert pre[:8 * cyc].sum() == b.n_bases                       # every base lands in exactly one (bin, cycle)
    assert pre[16 * cyc + abi.STATS_QUALHIST:][:128].sum() == b.n_bases
    passed = (res["seg_result"] == 0) & (np.arange(2)[None, :] < res["n_segments"][:, None])
    assert post[16 * cyc + abi.STATS_READS] == passed.sum() == cnt[abi.CNT_FILTER + 0]
    assert post[16 * cyc + abi.STATS_LENSUM] == res["seg_len"][passed].sum()
    assert cnt[abi.CNT_FILTER:abi.CNT_FILTER + 32].sum() == res["n_segments"].sum()
    # windows nest: segments inside the trim window inside the read
    lens = b.lens
    ok = res["n_segments"] > 0
    assert (res["trim_lo"][ok] >= 0).all() and ((res["trim_lo"] + res["trim_len"])[ok] <= lens[ok]).all()
    for k in range(2):
        m = res["n_segments"] > k
        assert (res["seg_lo"][m, k] >= res["trim_lo"][m]).all()
        assert ((res["seg_lo"][:, k] + res["seg_len"][:, k])[m] <= (res["trim_lo"] + res["trim_len"])[m]).all()
    # idempotence of the accumulators: a second pass doubles every counter
    g.process(b)
    assert np.array_equal(g.stats(0, cyc), 2 * pre) and np.array_equal(g.counters(), 2 * cnt)


def _device_batch(batch):
    import torch
    # FPL_EMULATE=1 (conftest.py): the emulated library's device memory is host memory
    dev = torch.device("cpu" if os.environ.get("FPL_EMULATE", "") not in ("", "0") else "cuda:0")
    pad = 256
    seq = torch.zeros(batch.seq.size + pad, dtype=torch.uint8, device=dev)
    qual = torch.zeros(batch.qual.size + pad, dtype=torch.uint8, device=dev)
    seq[:batch.seq.size] = torch.from_numpy(batch.seq).to(dev)
    qual[:batch.qual.size] = torch.from_numpy(batch.qual).to(dev)
    offs = torch.from_numpy(batch.offsets.astype(np.int64)).to(dev)
    lens = torch.from_numpy(batch.lens.astype(np.int32)).to(dev)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    return seq, qual, offs, lens


@pytest.mark.parametrize("tiling", [None, "2"])
def test_device_resident_entry_point(tiling, monkeypatch):
    """fpl_process_device (what bench.py's `value` leg and a GPU-resident caller use): same records, Stats blocks and
    counters as the oracle; with and without read tiling (without it the host only learns the longest read, from a
    device-side reduction on the side stream)."""
    if tiling:
        monkeypatch.setenv("FPL_TILE_MBASES", tiling)
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    batch = cases.ont_batch(21, n=900, mean=4000, p_chimera=0.05, p_polya=0.05)
    seq, qual, offs, lens = _device_batch(batch)
    g, o = gpu_engine(opt), OracleEngine(opt)
    for rep in range(2):      # back to back: the second call must not depend on the first one having drained
        g.process_device(seq.data_ptr(), qual.data_ptr(), offs.data_ptr(), lens.data_ptr(), batch.n_reads, batch.seq.size)
    g.sync()
    res = g.fetch_results(batch.n_reads)
    exp = o.process(batch)
    o.process(batch)
    compare_results(res, exp, "device")
    cyc = int(batch.lens.max())
    for w in (0, 1):
        compare_stats(g.stats(w, cyc), o.stats(w, cyc), f"device/stats{w}")
    compare_stats(g.counters(), o.counters(), "device/counters")
    # empty batch, then a negative length
    g.process_device(seq.data_ptr(), qual.data_ptr(), offs.data_ptr(), lens.data_ptr(), 0, 0)
    import torch
    bad = lens.clone()
    bad[3] = -5
    with pytest.raises(Exception):
        g.process_device(seq.data_ptr(), qual.data_ptr(), offs.data_ptr(), bad.data_ptr(), batch.n_reads, batch.seq.size)
    g.close()


def test_passing_segments_shorter_than_a_5mer():
    """--length_required 0 lets segments of 2-4 bases pass: the post-filter 5-mer table (pre minus removed) must not
    subtract the bases behind such a segment twice (found by tools/fuzz_gpu_vs_oracle.py)."""
    opt = Options(disable_adapter_trimming=True, cut_front=True, cut_tail=True, length_required=0, length_limit=500,
                  low_complexity_filter=True, complexity_threshold=60)
    check_against_oracle(opt, cases.adversarial_batch(825889), "short-segments")
    reads = [(b"ACGTACGTACGTACGTACGT"[:n], bytes([33 + 30]) * n) for n in (1, 2, 3, 4, 5, 6, 9)] * 3
    check_against_oracle(Options(disable_adapter_trimming=True, length_required=0), pack_reads(reads), "tiny-reads")


def test_random_option_sets_gpu():
    """Seeded random option sets x batches against the oracle (cases.random_case; tools/fuzz_gpu_vs_oracle.py runs the
    same generator for longer): every record, both Stats blocks, counters, and the --mask/--break lists."""
    import random
    from oracle_lib import compare_lists
    rng = random.Random(12)
    for i in range(60):
        opt, batch, what = cases.random_case(rng)
        g, o = gpu_engine(opt), OracleEngine(opt)
        try:
            compare_results(g.process(batch), o.process(batch), what)
            if opt.mask or opt.break_reads:
                compare_lists(g.segments(), o.segments(), what + "/segments")
                compare_lists(g.mask_regions(), o.mask_regions(), what + "/regions")
            cyc = max(1, int(batch.lens.max()))
            for w in (0, 1):
                compare_stats(g.stats(w, cyc), o.stats(w, cyc), f"{what}/stats{w}")
            compare_stats(g.counters(), o.counters(), what + "/counters")
        finally:
            g.close()
