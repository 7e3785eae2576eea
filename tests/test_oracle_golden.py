"""The oracle against the committed golden fixtures (tests/golden/, generated from the unmodified reference by
tests/golden/make_golden.py).  CPU only; does not need /root/reference."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from fastplong_b200 import PackedBatch, hostside, synth
from oracle_lib import OracleEngine, compare_results, compare_stats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_input():
    z = np.load(os.path.join(GOLDEN, "adversarial_input.npz"))
    return PackedBatch(z["seq"], z["qual"], z["offsets"], z["lens"])


@pytest.mark.parametrize("name", sorted(cases.OPTION_SETS))
def test_oracle_matches_reference_fixture(name):
    b = golden_input()
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    o = OracleEngine(cases.OPTION_SETS[name])
    res = o.process(b)
    compare_results(res, z["results"], name)
    cyc = int(z["cycles"])
    compare_stats(o.stats(0, cyc), z["pre"], name + "/pre")
    compare_stats(o.stats(1, cyc), z["post"], name + "/post")
    compare_stats(o.counters(), z["counters"], name + "/counters")


@pytest.mark.parametrize("name", ["c1_small", "cut_polyx", "loose"])
def test_oracle_reproduces_reference_binary_outputs(name):
    """Oracle records -> host FASTQ assembly must give the md5 of fastplong_ref's --out/--failed_out, and the
    report scalars must match the reference JSON: pins the driver order (seprocessor.cpp:180-329) end to end."""
    g = json.load(open(os.path.join(GOLDEN, f"binary_{name}.json")))
    opt = cases.OPTION_SETS[g["options"]]
    batch = synth.ont_like(g["n_reads"], g["mean_len"], g["seed"], **g["synth_kwargs"])
    fq = b"".join(b"@read%d len=%d\n%s\n+\n%s\n" % (i, len(s), s, q) for i, (s, q) in
                  enumerate(batch.read(i) for i in range(batch.n_reads)))
    if hashlib.md5(fq).hexdigest() != g["input_md5"]:
        pytest.skip("numpy generator produced a different synthetic input than when the fixture was made")
    o = OracleEngine(opt)
    res = o.process(batch)
    out, failed = hostside.emit_fastq(batch, hostside.default_names(batch), res)
    assert hashlib.md5(out).hexdigest() == g["out_md5"]
    assert hashlib.md5(failed).hexdigest() == g["failed_md5"]
    cyc = int(batch.lens.max())
    rep = hostside.report_summary(o.stats(0, cyc), o.stats(1, cyc), o.counters(), cyc)
    j = g["json"]
    for ours, key in ((rep["before"], "read_before_filtering"), (rep["after"], "read_after_filtering")):
        for f in ("total_reads", "total_bases", "q20_bases", "q30_bases", "total_cycles"):
            assert ours[f] == j[key][f], (key, f)
    assert rep["before"]["read_mean_length"] == j["summary"]["before_filtering"]["read_mean_length"]
    assert rep["after"]["read_mean_length"] == j["summary"]["after_filtering"]["read_mean_length"]
    for f, v in rep["filtering_result"].items():
        assert v == j["filtering_result"][f], f
    assert rep["adapter_trimmed_reads"] == j["adapter_cutting"]["adapter_trimmed_reads"]
    assert rep["adapter_trimmed_bases"] == j["adapter_cutting"]["adapter_trimmed_bases"]
    # adapter counts: the reference folds entries below 1% into "others" (src/filterresult.cpp:134-169)
    amap = hostside.adapter_count_map(o.counters(), opt.adapter_list())
    ref_counts = dict(j["adapter_cutting"]["read_adapter_counts"])
    others = ref_counts.pop("others", 0)
    for k, v in ref_counts.items():
        assert amap[k] == v, k
    assert sum(v for k, v in amap.items() if k not in ref_counts) == others


# ---- --mask / --break ----
from oracle_lib import compare_lists  # noqa: E402


def blocky_input():
    z = np.load(os.path.join(GOLDEN, "blocky_input.npz"))
    return PackedBatch(z["seq"], z["qual"], z["offsets"], z["lens"])


@pytest.mark.parametrize("name", sorted(cases.MASK_BREAK_SETS))
def test_oracle_matches_reference_fixture_mask_break(name):
    b = blocky_input()
    z = np.load(os.path.join(GOLDEN, f"ref_mb_{name}.npz"))
    o = OracleEngine(cases.MASK_BREAK_SETS[name])
    compare_results(o.process(b), z["results"], name)
    compare_lists(o.segments(), z["segments"], name + "/segments")
    compare_lists(o.mask_regions(), z["regions"], name + "/regions")
    cyc = int(z["cycles"])
    compare_stats(o.stats(0, cyc), z["pre"], name + "/pre")
    compare_stats(o.stats(1, cyc), z["post"], name + "/post")
    compare_stats(o.counters(), z["counters"], name + "/counters")


@pytest.mark.parametrize("name", ["mb_break", "mb_mask", "mb_both"])
def test_oracle_reproduces_reference_binary_outputs_mask_break(name):
    """Oracle records + output-read list + masked regions -> host FASTQ assembly == fastplong_ref -N / -b outputs
    ("r<k>-" names, masked bases, the failed-out rule of src/seprocessor.cpp:264-288)."""
    g = json.load(open(os.path.join(GOLDEN, f"binary_{name}.json")))
    opt = cases.MASK_BREAK_SETS[g["options"]]
    batch = cases.blocky_quality_batch(g["seed"], n=g["n_reads"])
    fq = b"".join(b"@read%d len=%d\n%s\n+\n%s\n" % (i, len(s), s, q) for i, (s, q) in
                  enumerate(batch.read(i) for i in range(batch.n_reads)))
    if hashlib.md5(fq).hexdigest() != g["input_md5"]:
        pytest.skip("numpy generator produced a different synthetic input than when the fixture was made")
    o = OracleEngine(opt)
    res = o.process(batch)
    out, failed = hostside.emit_fastq_ext(batch, hostside.default_names(batch), res, o.segments(), o.mask_regions())
    assert hashlib.md5(out).hexdigest() == g["out_md5"]
    assert hashlib.md5(failed).hexdigest() == g["failed_md5"]
    cyc = int(batch.lens.max())
    rep = hostside.report_summary(o.stats(0, cyc), o.stats(1, cyc), o.counters(), cyc)
    j = g["json"]
    for ours, key in ((rep["before"], "read_before_filtering"), (rep["after"], "read_after_filtering")):
        for f in ("total_reads", "total_bases", "q20_bases", "q30_bases", "total_cycles"):
            assert ours[f] == j[key][f], (key, f)
    for f, v in rep["filtering_result"].items():
        assert v == j["filtering_result"][f], f
