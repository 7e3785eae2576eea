"""The index arithmetic of k_cycle_stats' shared-memory counters (fastplong_b200/csrc/fpl_stats.cu), checked on the CPU:
the (lane, byte) -> word mapping is injective per bin, free of bank conflicts for every misalignment, and the flush maps
every word back to the cycle the byte came from.  (This checks the arithmetic the kernel's constants encode; that the
hardware sees the resulting address pattern as conflict-free — 0 conflicts, 1 wavefront per reduction — is measured by
tools/ubench_red.cu, profiles/r02k_ubench_red_ncu.csv.  The conflict wavefronts ncu reports inside the kernel come from the
reductions sharing the data pipe with the cp.async ring fills and other warps' loads)"""
import re
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "fastplong_b200", "csrc", "fpl_stats.cu")).read()


def define(name):
    m = re.search(r"#define\s+%s\s+(.+?)\s*(//.*)?$" % name, SRC, re.M)
    assert m, name
    return m.group(1)


ROWW = int(define("CS_ROWW"))
BINW = eval(define("CS_BINW").replace("CS_ROWW", str(ROWW)))


def word(lane, j, s):          # count1<S, J>: immediate 4*((m & 15)*ROWW + (m >> 4)) on top of pk_lane = base + 4*lane
    m = j + s
    return (m & 15) * ROWW + (m >> 4) + lane


def test_constants():
    assert ROWW == 33 and BINW % 32 == 0 and BINW >= 16 * ROWW


def test_mapping_is_injective_and_flush_inverts_it():
    for s in range(16):                       # S = 15 - (segment address & 15)
        seen = {}
        for lane in range(32):
            for j in range(16):
                w = word(lane, j, s)
                col = 16 * lane + j + s       # column of the tile: cycle = 512*tile - 15 + col
                assert 0 <= w < 16 * ROWW
                assert seen.setdefault(w, col) == col
                # the flush: word p -> column 16*(p % ROWW) + p // ROWW
                assert 16 * (w % ROWW) + w // ROWW == col


def test_no_bank_conflicts_for_any_misalignment_or_bin_mix():
    for s in range(16):
        for j in range(16):
            for bins in ([0] * 32, list(range(8)) * 4, [7, 0] * 16, [3, 5, 1, 6] * 8):
                banks = {(b * BINW + word(lane, j, s)) % 32 for lane, b in enumerate(bins)}
                assert len(banks) == 32, (s, j)


def test_kmer_code_permutation_is_a_bijection():
    """the 5-mer tables are indexed by pairs (b2, b1); the flush swaps the bits of every pair to get base2val's code"""
    codes = {((i & 0x155) << 1) | ((i >> 1) & 0x155) for i in range(1024)}
    assert codes == set(range(1024))
    cp = {"A": 0, "C": 1, "T": 2, "U": 2, "G": 3}          # (byte >> 1) & 3
    true = {"A": 0, "T": 1, "U": 1, "C": 2, "G": 3}        # Stats::base2val
    for ch, c in cp.items():
        assert (ord(ch) >> 1) & 3 == c and (((c & 1) << 1) | (c >> 1)) == true[ch]
