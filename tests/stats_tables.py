"""TEST INFRASTRUCTURE (checker, never the product path): the Stats block of a whole batch restated with plain torch ops.

`Stats::statRead` (/root/reference/src/stats.cpp:265-375) over a list of segments is, table by table, a histogram of the
segments' bytes: contents / quality sums keyed by (base & 7, cycle), the quality histogram, the 5-mer table
(closed form: every position i >= 4 whose five bases are all in ACGTU), and per segment the median quality character.
Here each of them is one `torch.bincount` over byte chunks of the packed buffers, so the same code checks a 1 500-read
batch against the C oracle on the CPU (tests/test_stats_tables.py pins this restatement to the oracle) and the
1 M-read, 32 GB batch of bench.py on the GPU (`parity_full_scale`), where the oracle would take an hour.  It shares no code
with the library's kernels (no tiling by cycle, no packed counters, no "post = pre - removed" derivation).

Layout of the result = FPL_STATS_WORDS(cap) of include/fplgpu.h, as fastplong_b200.abi describes it.
"""
import numpy as np

from fastplong_b200 import abi

CHUNK = 1 << 26


def _code_table(torch, dev):
    t = torch.full((256,), -1, dtype=torch.int64, device=dev)          # Stats::base2val, src/stats.cpp:411-425
    for ch, v in ((65, 0), (84, 1), (85, 1), (67, 2), (71, 3)):
        t[ch] = v
    return t


def segment_medians(torch, hist):
    """hist [n, 128] (int64), per-segment quality histogram -> the median character of src/stats.cpp:351-361: the smallest m
    with sum_{c <= m} hist[c] > len >> 1 (0 for an empty segment, which the caller ignores)."""
    cum = hist.cumsum(dim=1)
    half = (cum[:, -1] >> 1)[:, None]
    return (cum <= half).sum(dim=1)


def stats_block(torch, seq, qual, starts, lens, cap, chunk=CHUNK, deadline=None):
    """seq, qual: uint8 tensors (any device); starts / lens: int64 numpy arrays, one entry per segment handed to statRead,
    in any order, segments not overlapping.  Returns (block, medians): block = int64 numpy vector of abi.stats_words(cap)
    words, medians = uint8 numpy array per segment (0 for empty segments).  deadline: time.time() value beyond which the
    walk raises TimeoutError (bench.py bounds the check; every chunk synchronises, so the clock is meaningful)."""
    import time
    dev = seq.device
    starts = np.asarray(starts, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int64)
    n_all = len(lens)
    block = np.zeros(abi.stats_words(cap), dtype=np.int64)
    tail = block[16 * cap:]
    tail[abi.STATS_READS] = n_all
    tail[abi.STATS_LENSUM] = int(lens.sum())
    medians = np.zeros(n_all, dtype=np.uint8)
    keep = np.nonzero(lens > 0)[0]
    if len(keep) == 0:
        return block, medians
    order = keep[np.argsort(starts[keep], kind="stable")]
    st_h, ln_h = starts[order], lens[order]
    assert int(ln_h.max()) <= cap, "a segment is longer than the block's cycle capacity"
    assert (st_h[1:] >= (st_h + ln_h)[:-1]).all(), "segments overlap"
    st = torch.from_numpy(st_h).to(dev)
    en = torch.from_numpy(st_h + ln_h).to(dev)
    code_of = _code_table(torch, dev)
    content = torch.zeros(8 * cap, dtype=torch.int64, device=dev)
    qsum = torch.zeros(8 * cap, dtype=torch.float64, device=dev)       # exact: every sum stays far below 2^53
    qhist = torch.zeros(128, dtype=torch.int64, device=dev)
    kmer = torch.zeros(1024, dtype=torch.int64, device=dev)
    seg_hist = torch.zeros((len(order), 128), dtype=torch.int32, device=dev)
    first, last = int(st_h[0]), int((st_h + ln_h).max())
    for lo in range(first, last, chunk):
        if deadline is not None and time.time() > deadline:
            raise TimeoutError("stats_block: time budget used up at byte %d of [%d, %d)" % (lo, first, last))
        hi = min(last, lo + chunk)
        lo_e = max(0, lo - 4)                                           # four bytes of run-in for the 5-mers
        idx = torch.arange(lo_e, hi, dtype=torch.int64, device=dev)
        r = torch.searchsorted(st, idx, right=True) - 1
        rc = r.clamp(min=0)
        s0 = st[rc]
        inside = (r >= 0) & (idx < en[rc])
        own = inside & (idx >= lo)
        if not bool(own.any()):
            continue
        cyc = idx - s0
        s = seq[lo_e:hi].to(torch.int64)
        q = qual[lo_e:hi].to(torch.int64)
        key = ((s & 7) * cap + cyc)[own]
        qo = q[own]
        content += torch.bincount(key, minlength=8 * cap)
        qsum += torch.bincount(key, weights=(qo - 33).to(torch.float64), minlength=8 * cap)
        qhist += torch.bincount(qo, minlength=128)[:128]
        # per-segment histograms of the segments this chunk touches
        ro = rc[own]
        r0, r1 = int(ro.min()), int(ro.max())
        h = torch.bincount((ro - r0) * 128 + qo, minlength=(r1 - r0 + 1) * 128)
        seg_hist[r0:r1 + 1] += h.reshape(-1, 128).to(torch.int32)
        # 5-mers ending at i: i >= 4 inside its segment and the five codes valid; first base most significant
        code = code_of[s]
        m = code.numel()
        if m >= 5:
            good = code >= 0
            ok = own[4:] & (cyc[4:] >= 4)
            kk = torch.zeros(m - 4, dtype=torch.int64, device=dev)
            for j in range(5):                                          # j = 0: the base four positions back
                ok = ok & good[j:m - 4 + j]
                kk = kk * 4 + code[j:m - 4 + j].clamp(min=0)
            kmer += torch.bincount(kk[ok], minlength=1024)
        del idx, r, rc, s0, inside, own, cyc, s, q, key, qo, ro, h, code
    block[:8 * cap] = content.cpu().numpy()
    block[8 * cap:16 * cap] = np.rint(qsum.cpu().numpy()).astype(np.int64)
    tail[abi.STATS_KMER:abi.STATS_KMER + 1024] = kmer.cpu().numpy()
    tail[abi.STATS_QUALHIST:abi.STATS_QUALHIST + 128] = qhist.cpu().numpy()
    med = segment_medians(torch, seg_hist.to(torch.int64)).cpu().numpy()
    medians[order] = med.astype(np.uint8)
    tail[abi.STATS_MEDHIST:abi.STATS_MEDHIST + 128] = np.bincount(med, minlength=128)[:128]
    tail[abi.STATS_MEDBASES:abi.STATS_MEDBASES + 128] = np.bincount(med, weights=ln_h.astype(np.float64), minlength=128)[:128].astype(np.int64)
    return block, medians


def passing_segments(res, offsets):
    """(read index, k) of every segment the post-filter Stats see (result code PASS_FILTER, src/seprocessor.cpp:270-276):
    returns (read [m], k [m], starts [m] absolute byte positions, lens [m])."""
    passed = (res["seg_result"] == abi.PASS_FILTER) & (np.arange(2)[None, :] < res["n_segments"][:, None])
    rd, k = np.nonzero(passed)
    starts = np.asarray(offsets, dtype=np.int64)[rd] + res["seg_lo"][rd, k].astype(np.int64)
    return rd, k, starts, res["seg_len"][rd, k].astype(np.int64)
