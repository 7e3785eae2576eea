"""--mask / --break on the device against the oracle: records, the full list of output reads, the
masked regions, both Stats blocks and the counters."""
import pytest

import cases
from oracle_lib import OracleEngine, compare_lists, compare_results, compare_stats

pytestmark = pytest.mark.gpu


def check(opt, batch, what):
    from fastplong_b200.binding import Engine
    g, o = Engine(opt), OracleEngine(opt)
    compare_results(g.process(batch), o.process(batch), what)
    compare_lists(g.segments(), o.segments(), what + "/segments")
    compare_lists(g.mask_regions(), o.mask_regions(), what + "/mask regions")
    cyc = max(1, int(batch.lens.max()))
    for w in (0, 1):
        compare_stats(g.stats(w, cyc), o.stats(w, cyc), f"{what}/stats{w}")
    compare_stats(g.counters(), o.counters(), what + "/counters")


@pytest.mark.parametrize("name", sorted(cases.MASK_BREAK_SETS))
@pytest.mark.parametrize("seed", [1, 2])
def test_mask_break_blocky_quality(name, seed):
    check(cases.MASK_BREAK_SETS[name], cases.blocky_quality_batch(70 + seed), f"{name}/blocky{seed}")


@pytest.mark.parametrize("name", sorted(cases.MASK_BREAK_SETS))
def test_mask_break_adversarial(name):
    check(cases.MASK_BREAK_SETS[name], cases.adversarial_batch(5), f"{name}/adv")
