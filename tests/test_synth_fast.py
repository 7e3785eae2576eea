"""The vectorised generator bench.py draws its workloads from (fastplong_b200/synth_fast.py), on the CPU."""
import numpy as np

from fastplong_b200 import Options, synth, synth_fast
from fastplong_b200.pack import SLOT_ALIGN
from oracle_lib import OracleEngine


def test_layout_alphabet_and_determinism():
    a = synth_fast.ont_like_fast(300, 3000, 11)
    b = synth_fast.ont_like_fast(300, 3000, 11)
    c = synth_fast.ont_like_fast(300, 3000, 12)
    assert np.array_equal(a.seq, b.seq) and np.array_equal(a.qual, b.qual) and np.array_equal(a.lens, b.lens)
    assert not np.array_equal(a.seq[:1000], c.seq[:1000])
    assert (a.offsets % SLOT_ALIGN == 0).all() and (a.lens >= 200).all()
    assert a.offsets[-1] + a.lens[-1] <= a.n_bytes
    for i in range(0, 300, 17):
        s, q = a.read(i)
        assert set(s) <= set(b"ACGTN")
        qa = np.frombuffer(q, dtype=np.uint8)
        assert qa.min() >= 34 and qa.max() <= 83 and (qa[:20] <= 44).all()     # first 20 qualities degraded to <= 11


def test_planted_structure_is_what_the_pipeline_finds():
    """~80 % of the reads carry a noisy start adapter, ~70 % an end adapter: the oracle trims most of them."""
    b = synth_fast.ont_like_fast(400, 4000, 5, p_chimera=0.05)
    opt = Options(start_adapter=synth.ADAPTER_START, end_adapter=synth.ADAPTER_END)
    res = OracleEngine(opt).process(b)
    trimmed = (res["adapter_trimmed_bases"] > 0).mean()
    assert 0.7 < trimmed <= 1.0
    assert 0.01 < ((res["flags"] & 4) != 0).mean() < 0.2          # chimeras split by the middle-adapter scan


def test_device_batch_slice_roundtrip():
    d = synth_fast.ont_like_device(64, 2000, 3, "cpu")
    whole = synth_fast.ont_like_fast(64, 2000, 3)
    part = d.to_host(10, 30)
    assert part.n_reads == 20 and part.offsets[0] == 0
    for i in range(20):
        assert part.read(i) == whole.read(10 + i)
    assert d.to_host(5, 5).n_reads == 0


def test_shared_plan_gives_equal_work_and_different_bases():
    plan = synth_fast.read_plan(100, 3000, 9)
    a = synth_fast.ont_like_device(100, 3000, 100, "cpu", plan=plan)
    b = synth_fast.ont_like_device(100, 3000, 200, "cpu", plan=plan)
    assert np.array_equal(a.lens, b.lens) and a.n_bases == b.n_bases
    assert not np.array_equal(a.seq.numpy()[:4000], b.seq.numpy()[:4000])


def test_hifi_like_options():
    fa = ["ACGTACGTACGTACGTACGTAC", "TTGACCATGGACCATGACCAGTTA"]
    b = synth_fast.ont_like_fast(200, 3000, 2, q_mean=33.0, q_sd=6.0, q_clip=60, p_polya=0.3, planted=fa, p_planted=0.5)
    tails = sum(1 for i in range(200) if b.read(i)[0][-15:] in (b"A" * 15, b"T" * 15))
    assert 30 < tails < 100
    assert max(np.frombuffer(b.read(0)[1], dtype=np.uint8)) <= 93
