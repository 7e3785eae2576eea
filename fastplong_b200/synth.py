"""Seeded synthetic long-read generator — the bench/test workloads of BASELINE.json.

Everything is derived from one numpy Generator seed, so a workload is reproducible on the GPU box without
shipping data.  `ont_like` builds a packed batch directly (vectorised for the 1 Gbase tiles bench.py uses);
`adversarial_reads` is a small hand-built set that exercises the quirk list.
"""
import numpy as np

from .pack import PackedBatch, slot_offsets
from .options import reverse_complement

ADAPTER_START = "AATGTACTTCGTTCAGTTACGTATTGCTAA"  # 30 bp
ADAPTER_END = reverse_complement(ADAPTER_START)
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _noisy(rng, s, rate=0.08):
    """per-base deletion / insertion / substitution, total rate `rate`"""
    out = bytearray()
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(int(BASES[rng.integers(4)]))
            out.append(ch)
        elif r < rate:
            out.append(int(BASES[rng.integers(4)]))
        else:
            out.append(ch)
    return bytes(out)


def _homopolymer(rng, lo, hi):
    n = int(rng.integers(lo, hi + 1))
    return bytes([int(BASES[rng.integers(4)])]) * n


def ont_like(n_reads, mean_len, seed, *, q_mean=18.0, q_sd=7.0, q_clip=50, p_start=0.8, p_end=0.7,
             p_chimera=0.01, p_n=0.001, p_polya=0.0, adapter_start=ADAPTER_START, adapter_end=ADAPTER_END,
             min_len=200, planted=(), p_planted=0.0):
    """n_reads reads, body length max(min_len, Gamma(2, mean_len/2)); see SURVEY Appendix C."""
    rng = np.random.default_rng(seed)
    body = np.maximum(min_len, rng.gamma(2.0, mean_len / 2.0, size=n_reads).astype(np.int64))
    a_s, a_e = adapter_start.encode(), adapter_end.encode()
    heads, tails, mids, midpos = [], [], [], []
    for i in range(n_reads):
        h = t = m = b""
        if rng.random() < p_start:
            h = _homopolymer(rng, 0, 24) + _noisy(rng, a_s)
        if planted and rng.random() < p_planted:
            h = _noisy(rng, planted[int(rng.integers(len(planted)))].encode()) + h
        if rng.random() < p_end:
            t = _noisy(rng, a_e) + _homopolymer(rng, 0, 14)
        if p_polya and rng.random() < p_polya:
            t = t + bytes([int(BASES[[0, 3][int(rng.integers(2))]])]) * int(rng.integers(15, 61))
        if rng.random() < p_chimera:
            m = _noisy(rng, a_e) + _noisy(rng, a_s)
        heads.append(h); tails.append(t); mids.append(m)
        midpos.append(int(rng.integers(0, body[i])) if m else 0)
    lens = body + np.array([len(h) + len(t) + len(m) for h, t, m in zip(heads, tails, mids)], dtype=np.int64)
    offsets, total = slot_offsets(lens)
    seq = BASES[rng.integers(0, 4, size=total, dtype=np.uint8)]
    q = np.rint(rng.normal(q_mean, q_sd, size=total)).clip(1, q_clip).astype(np.uint8) + 33
    if p_n > 0:
        nmask = rng.random(total, dtype=np.float32) < p_n
        seq[nmask] = ord("N")
    for i in range(n_reads):
        o, L = int(offsets[i]), int(lens[i])
        h, t, m = heads[i], tails[i], mids[i]
        if h:
            seq[o:o + len(h)] = np.frombuffer(h, dtype=np.uint8)
        if t:
            seq[o + L - len(t):o + L] = np.frombuffer(t, dtype=np.uint8)
        if m:
            p = o + len(h) + midpos[i]
            p = min(p, o + L - len(t) - len(m))
            seq[p:p + len(m)] = np.frombuffer(m, dtype=np.uint8)
    # first 20 qualities degraded: min(q, U[1,11])
    k = np.minimum(20, lens)
    idx = (offsets[:, None] + np.arange(20)[None, :])
    valid = np.arange(20)[None, :] < k[:, None]
    low = (rng.integers(1, 12, size=idx.shape).astype(np.uint8) + 33)
    flat = idx[valid]
    q[flat] = np.minimum(q[flat], low[valid])
    return PackedBatch(seq, q, offsets, lens.astype(np.int32))


def to_fastq(batch, path, name_prefix="read"):
    """Write the batch as plain FASTQ (name line `@read<i> len=<L>`, plus line `+`)."""
    with open(path, "wb") as f:
        for i in range(batch.n_reads):
            s, q = batch.read(i)
            f.write(b"@%s%d len=%d\n" % (name_prefix.encode(), i, len(s)))
            f.write(s + b"\n+\n" + q + b"\n")


def adversarial_reads(seed, n_random=200, adapter_start=ADAPTER_START, adapter_end=ADAPTER_END):
    """Small reads built to hit the quirks of SURVEY A.10 plus random short/odd reads."""
    rng = np.random.default_rng(seed)
    a_s, a_e = adapter_start.encode(), adapter_end.encode()

    def rb(n):
        return BASES[rng.integers(0, 4, size=n)].tobytes()

    def rq(n, mean=18, sd=7):
        return (np.rint(rng.normal(mean, sd, size=n)).clip(1, 50).astype(np.uint8) + 33).tobytes()

    reads = []

    def add(s, q=None, **kw):
        reads.append((s, q if q is not None else rq(len(s), **kw)))

    add(b"")                                    # empty read
    add(b"A"); add(b"N"); add(b"ACGTN" * 3)     # tiny reads
    add(b"A" * 30)                              # polyX runs off the front
    add(b"N" * 40)
    add(rb(300) + b"A" * 25)                    # polyA tail
    add(rb(300) + b"ATTTTTTTTTTTTTTTTTTTTTTT")  # polyT with junk
    add(rb(100) + b"N" * 12 + b"G" * 14)
    add(rb(500), rq(500, mean=5, sd=2))         # all low quality
    add(rb(500), rq(500, mean=40, sd=2))        # all high quality
    add(rb(15)); add(rb(16)); add(rb(17)); add(rb(29)); add(rb(30)); add(rb(31)); add(rb(45)); add(rb(46))
    add(a_s); add(a_e); add(a_s + a_e); add(a_s[5:] + rb(50)); add(rb(50) + a_e[:20])
    add(a_s + rb(400) + a_e)                    # both adapters exact
    add(rb(10) + a_s + rb(400) + a_e + rb(5))
    add(a_s[10:] + rb(300) + a_e[:18])          # partial adapters flush with the ends
    add(a_s[14:] + rb(300) + a_e[:16])
    add(rb(300) + a_e[:17])
    add(rb(200) + a_e + a_s + rb(250))          # chimera
    add(rb(200) + a_s + rb(250))                # middle start adapter only
    add(rb(5) + a_e + rb(300))                  # adapter near the start -> empty left segment
    add(rb(300) + a_s + rb(3))
    add(a_s + a_s + rb(100)); add(rb(100) + a_e + a_e)
    add(b"AC" * 200); add(b"A" * 200 + b"C" * 200)  # low complexity
    add(b"N" * 5 + rb(300) + b"N" * 7)
    add(rb(150) + b"N" * 30 + rb(150))
    add(b"acgtnACGTNUuXx" * 20)                 # odd alphabet, lower case, U
    add(rb(300), bytes([33]) * 300)             # minimum quality
    add(rb(300), bytes([126]) * 300)            # maximum printable quality
    add(rb(64)); add(rb(199)); add(rb(200)); add(rb(201)); add(rb(215)); add(rb(216)); add(rb(217))
    for _ in range(n_random):
        L = int(rng.integers(0, 700))
        s = bytearray(rb(L))
        kind = rng.integers(0, 8)
        if kind == 0 and L > 40:
            piece = _noisy(rng, a_s, 0.1)
            s[:len(piece)] = piece[:L]
        elif kind == 1 and L > 40:
            piece = _noisy(rng, a_e, 0.1)
            s[L - len(piece):] = piece[-L:] if len(piece) <= L else piece[:L]
        elif kind == 2 and L > 80:
            piece = _noisy(rng, a_s, 0.05)
            p = int(rng.integers(0, L - len(piece)))
            s[p:p + len(piece)] = piece
        elif kind == 3 and L > 30:
            k = int(rng.integers(8, 30))
            base = int(BASES[rng.integers(4)])
            tail = bytearray([base]) * k
            for p in rng.integers(0, k, size=int(rng.integers(0, 3))):
                tail[int(p)] = int(BASES[rng.integers(4)])
            s[L - k:] = tail
        elif kind == 4:
            for p in rng.integers(0, max(1, L), size=max(1, L // 20)):
                if L:
                    s[int(p)] = ord("N")
        elif kind == 5 and L > 20:
            k = int(rng.integers(8, 20))
            s[L - k:] = bytes([int(BASES[rng.integers(4)])]) * k
        elif kind == 6 and L > 60:
            k = int(rng.integers(4, 29))
            s[:30 - k] = a_s[k:]
        elif kind == 7 and L > 60:
            k = int(rng.integers(16, 30))
            s[L - k:] = a_e[:k]
        mean = float(rng.choice([6, 12, 18, 30]))
        add(bytes(s), rq(L, mean=mean))
    return reads
