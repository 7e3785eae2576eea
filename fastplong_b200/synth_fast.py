"""Vectorised twin of synth.ont_like for workloads of BASELINE.json's size: one million DISTINCT
reads take seconds instead of the minutes the per-read Python loop of synth.ont_like needs.

The per-read structure (which read gets which noisy adapter copy, homopolymer, chimera insert, poly-A tail) is drawn on
the host with numpy, a whole batch at a time; the bulk bytes (random bodies, qualities, N bases) are drawn by torch
directly on the target device (cuda for bench.py, cpu for the reference arm and the tests) and the short per-read
pieces are scattered into them there.  Same distributions as synth.ont_like, not the same random stream.
"""
from dataclasses import dataclass

import numpy as np
import torch

from .pack import PackedBatch, slot_offsets, TAIL_PAD
from .synth import ADAPTER_END, ADAPTER_START

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _noisy_batch(rng, adapter, n, rate=0.08):
    """n independent noisy copies of `adapter` (per-base deletion / insertion / substitution, total rate `rate`):
    chars [n, 2*len] left-justified, lens [n]."""
    a = np.frombuffer(adapter, dtype=np.uint8)
    m = len(a)
    if m == 0 or n == 0:
        return np.zeros((n, 1), dtype=np.uint8), np.zeros(n, dtype=np.int64)
    r = rng.random((n, m))
    rnd = BASES[rng.integers(0, 4, size=(n, m))]
    dele = r < rate / 3
    ins = (r >= rate / 3) & (r < 2 * rate / 3)
    sub = (r >= 2 * rate / 3) & (r < rate)
    emit = np.where(dele, 0, np.where(ins, 2, 1)).astype(np.int64)
    pos = np.cumsum(emit, axis=1) - emit
    out = np.zeros((n, 2 * m), dtype=np.uint8)
    rows = np.broadcast_to(np.arange(n)[:, None], pos.shape)
    first = np.where(ins | sub, rnd, a[None, :])
    m1 = emit >= 1
    out[rows[m1], pos[m1]] = first[m1]
    m2 = emit == 2
    out[rows[m2], pos[m2] + 1] = np.broadcast_to(a[None, :], pos.shape)[m2]
    return out, emit.sum(axis=1)


def _homopolymer_batch(rng, n, lo, hi):
    lens = rng.integers(lo, hi + 1, size=n).astype(np.int64)
    base = BASES[rng.integers(0, 4, size=n)]
    return np.broadcast_to(base[:, None], (n, max(hi, 1))).copy(), lens


def _concat(pieces):
    """[(chars [n, w_k], lens [n])] -> (chars [n, sum w_k] left-justified concatenation, lens [n])"""
    n = pieces[0][0].shape[0]
    width = sum(p[0].shape[1] for p in pieces)
    out = np.zeros((n, width), dtype=np.uint8)
    start = np.zeros(n, dtype=np.int64)
    for chars, lens in pieces:
        w = chars.shape[1]
        cols = np.arange(w)[None, :]
        mask = cols < lens[:, None]
        rows = np.broadcast_to(np.arange(n)[:, None], mask.shape)
        out[rows[mask], (start[:, None] + cols)[mask]] = chars[mask]
        start = start + lens
    return out, start


@dataclass
class DeviceBatch:
    """A packed batch whose byte buffers live on a torch device (fpl_batch layout, include/fplgpu.h)."""
    seq: torch.Tensor        # uint8 [n_bytes]
    qual: torch.Tensor
    offsets: np.ndarray      # int64 [n_reads] (host)
    lens: np.ndarray         # int32 [n_reads] (host)

    @property
    def n_reads(self):
        return int(self.lens.shape[0])

    @property
    def n_bases(self):
        return int(self.lens.sum(dtype=np.int64))

    @property
    def n_bytes(self):
        return int(self.seq.shape[0])

    def to_host(self, lo=0, hi=None):
        """Reads [lo, hi) as a PackedBatch in host memory (offsets rebased to the slice)."""
        hi = self.n_reads if hi is None else hi
        if hi <= lo:
            return PackedBatch(np.zeros(TAIL_PAD, np.uint8), np.zeros(TAIL_PAD, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int32))
        b0 = int(self.offsets[lo])
        b1 = int(self.offsets[hi]) if hi < self.n_reads else self.n_bytes - TAIL_PAD
        seq = np.zeros(b1 - b0 + TAIL_PAD, dtype=np.uint8)
        qual = np.zeros_like(seq)
        seq[:b1 - b0] = self.seq[b0:b1].cpu().numpy()
        qual[:b1 - b0] = self.qual[b0:b1].cpu().numpy()
        return PackedBatch(seq, qual, (self.offsets[lo:hi] - b0).copy(), self.lens[lo:hi].copy())


def read_plan(n_reads, mean_len, seed, *, p_start=0.8, p_end=0.7, p_chimera=0.01, p_polya=0.0,
              adapter_start=ADAPTER_START, adapter_end=ADAPTER_END, min_len=200, planted=(), p_planted=0.0):
    """The host half: per-read lengths and the short pieces planted at the ends / in the middle.
    Returns dict(lens, heads=(chars, lens), tails=(chars, lens), mids=(chars, lens), midpos)."""
    rng = np.random.default_rng(seed)
    n = n_reads
    body = np.maximum(min_len, rng.gamma(2.0, mean_len / 2.0, size=n).astype(np.int64))
    a_s, a_e = adapter_start.encode(), adapter_end.encode()
    # head = [noisy planted adapter] + homopolymer(0..24) + noisy(start adapter), with probability p_start
    hp, hpl = _homopolymer_batch(rng, n, 0, 24)
    ns, nsl = _noisy_batch(rng, a_s, n)
    has = rng.random(n) < p_start
    pieces = []
    if planted and p_planted > 0:
        which = rng.integers(0, len(planted), size=n)
        hasp = rng.random(n) < p_planted
        width = 2 * max(len(p) for p in planted)
        pc = np.zeros((n, width), dtype=np.uint8)
        pl = np.zeros(n, dtype=np.int64)
        for k, pstr in enumerate(planted):
            sel = np.nonzero(hasp & (which == k))[0]
            c, l = _noisy_batch(rng, pstr.encode(), len(sel))
            pc[sel, :c.shape[1]] = c
            pl[sel] = l
        pieces.append((pc, pl))
    pieces += [(hp, np.where(has, hpl, 0)), (ns, np.where(has, nsl, 0))]
    heads = _concat(pieces)
    # tail = noisy(end adapter) + homopolymer(0..14) with probability p_end, then an optional poly-A/T tail
    ne, nel = _noisy_batch(rng, a_e, n)
    tp, tpl = _homopolymer_batch(rng, n, 0, 14)
    hae = rng.random(n) < p_end
    pieces = [(ne, np.where(hae, nel, 0)), (tp, np.where(hae, tpl, 0))]
    if p_polya > 0:
        pa = rng.random(n) < p_polya
        base = np.frombuffer(b"AT", dtype=np.uint8)[rng.integers(0, 2, size=n)]
        pal = np.where(pa, rng.integers(15, 61, size=n), 0).astype(np.int64)
        pieces.append((np.broadcast_to(base[:, None], (n, 60)).copy(), pal))
    tails = _concat(pieces)
    # chimera: noisy(end) + noisy(start) somewhere in the body, with probability p_chimera
    chim = rng.random(n) < p_chimera
    c1, l1 = _noisy_batch(rng, a_e, n)
    c2, l2 = _noisy_batch(rng, a_s, n)
    mids = _concat([(c1, np.where(chim, l1, 0)), (c2, np.where(chim, l2, 0))])
    midpos = (rng.random(n) * body).astype(np.int64)
    lens = body + heads[1] + tails[1] + mids[1]
    return dict(lens=lens, heads=heads, tails=tails, mids=mids, midpos=midpos)


def ont_like_device(n_reads, mean_len, seed, device, *, q_mean=18.0, q_sd=7.0, q_clip=50, p_n=0.001, plan=None, **kw):
    """n_reads DISTINCT reads in the packed layout, byte buffers on `device` (a torch device, cuda or cpu).
    `plan` (from read_plan) lets several ranks share one length distribution: equal work per GPU (weak scaling)."""
    dev = torch.device(device)
    if plan is None:
        plan = read_plan(n_reads, mean_len, seed, **kw)
    lens = plan["lens"]
    offsets, total = slot_offsets(lens)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) * 2654435761 % (2 ** 63))
    seq = torch.empty(total, dtype=torch.uint8, device=dev)
    qual = torch.empty(total, dtype=torch.uint8, device=dev)
    CH = 1 << 28
    for lo in range(0, total, CH):
        n = min(CH, total - lo)
        c = torch.randint(0, 4, (n,), generator=g, device=dev, dtype=torch.uint8)
        s = 65 + 2 * c + 2 * (c == 2).to(torch.uint8) + 13 * (c == 3).to(torch.uint8)     # A C G T
        if p_n > 0:
            s = torch.where(torch.rand(n, generator=g, device=dev) < p_n, torch.full_like(s, 78), s)
        seq[lo:lo + n] = s
        q = torch.randn(n, generator=g, device=dev).mul_(q_sd).add_(q_mean).round_().clamp_(1, q_clip).add_(33)
        qual[lo:lo + n] = q.to(torch.uint8)
        del c, s, q
    d_off = torch.from_numpy(offsets).to(dev)
    d_len = torch.from_numpy(lens).to(dev)

    def scatter(chars, plen, dst):
        """chars [n, w] (host), plen [n], dst [n] absolute byte positions (device int64)"""
        w = chars.shape[1]
        RB = 1 << 18
        for lo in range(0, chars.shape[0], RB):
            hi = min(lo + RB, chars.shape[0])
            pl = torch.from_numpy(plen[lo:hi]).to(dev)
            if int(pl.max()) == 0:
                continue
            ch = torch.from_numpy(np.ascontiguousarray(chars[lo:hi])).to(dev)
            cols = torch.arange(w, device=dev)[None, :]
            mask = cols < pl[:, None]
            idx = (dst[lo:hi, None] + cols)[mask]
            seq[idx] = ch[mask]

    hc, hl = plan["heads"]
    tc, tl = plan["tails"]
    mc, ml = plan["mids"]
    d_hl, d_tl, d_ml = (torch.from_numpy(x).to(dev) for x in (hl, tl, ml))
    scatter(hc, hl, d_off)
    scatter(tc, tl, d_off + d_len - d_tl)
    mid_dst = torch.minimum(d_off + d_hl + torch.from_numpy(plan["midpos"]).to(dev), d_off + d_len - d_tl - d_ml)
    scatter(mc, ml, mid_dst)
    # first 20 qualities degraded: min(q, U[1,11]) (every read is longer than 20)
    RB = 1 << 20
    for lo in range(0, n_reads, RB):
        hi = min(lo + RB, n_reads)
        idx = (d_off[lo:hi, None] + torch.arange(20, device=dev)[None, :]).reshape(-1)
        low = (torch.randint(1, 12, (idx.numel(),), generator=g, device=dev, dtype=torch.uint8) + 33)
        qual[idx] = torch.minimum(qual[idx], low)
    return DeviceBatch(seq, qual, offsets, lens.astype(np.int32))


def ont_like_fast(n_reads, mean_len, seed, **kw):
    """Host PackedBatch from the vectorised generator (the reference arm's FASTQ sample, CPU tests)."""
    b = ont_like_device(n_reads, mean_len, seed, "cpu", **kw)
    return PackedBatch(b.seq.numpy(), b.qual.numpy(), b.offsets, b.lens)
