"""Packed read batches: the HBM data layout of the hot path (include/fplgpu.h: fpl_batch).

Two byte buffers (sequence, quality) with identical layout; read i occupies [offsets[i], offsets[i]+lens[i]).
Slots start on SLOT_ALIGN-byte boundaries so that a warp's 16-byte vector loads and 1-D TMA bulk copies of
a read chunk are aligned; the pad bytes are never interpreted.  TAIL_PAD bytes after the last slot let kernels
over-read whole vectors.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from .abi import FplBatch

SLOT_ALIGN = 128
TAIL_PAD = 256


@dataclass
class PackedBatch:
    seq: np.ndarray      # uint8 [n_bytes]
    qual: np.ndarray     # uint8 [n_bytes]
    offsets: np.ndarray  # int64 [n_reads]
    lens: np.ndarray     # int32 [n_reads]

    @property
    def n_reads(self):
        return int(self.lens.shape[0])

    @property
    def n_bases(self):
        return int(self.lens.sum(dtype=np.int64))

    @property
    def n_bytes(self):
        return int(self.seq.shape[0])

    def to_abi(self):
        return FplBatch(self.seq.ctypes.data, self.qual.ctypes.data, self.offsets.ctypes.data,
                        self.lens.ctypes.data, self.n_reads, self.n_bytes)

    def read(self, i):
        o, n = int(self.offsets[i]), int(self.lens[i])
        return self.seq[o:o + n].tobytes(), self.qual[o:o + n].tobytes()

    def slice(self, lo, hi):
        """Reads [lo, hi) as views of the same buffers, offsets rebased to the slice's first byte: a rank that
        uploads its shard moves its own bytes only (slots are contiguous and in increasing order)."""
        if hi <= lo:
            return PackedBatch(self.seq[:TAIL_PAD], self.qual[:TAIL_PAD], self.offsets[:0].copy(), self.lens[:0].copy())
        b0 = int(self.offsets[lo])
        b1 = int(self.offsets[hi]) if hi < self.n_reads else self.n_bytes - TAIL_PAD
        b1 = max(b1, int(self.offsets[hi - 1]) + int(self.lens[hi - 1]))
        return PackedBatch(self.seq[b0:b1 + TAIL_PAD], self.qual[b0:b1 + TAIL_PAD], self.offsets[lo:hi] - b0,
                           self.lens[lo:hi].copy())


def slot_offsets(lens, align=SLOT_ALIGN):
    lens = np.asarray(lens, dtype=np.int64)
    slots = (lens + align - 1) // align * align
    offsets = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        np.cumsum(slots[:-1], out=offsets[1:])
    total = int(slots.sum()) + TAIL_PAD
    return offsets, total


def pack_reads(reads, align=SLOT_ALIGN):
    """reads: iterable of (seq_bytes, qual_bytes)."""
    reads = list(reads)
    lens = np.array([len(s) for s, _ in reads], dtype=np.int32)
    offsets, total = slot_offsets(lens, align)
    seq = np.zeros(total, dtype=np.uint8)
    qual = np.zeros(total, dtype=np.uint8)
    for (s, q), o, n in zip(reads, offsets, lens):
        assert len(q) == n
        seq[o:o + n] = np.frombuffer(s, dtype=np.uint8)
        qual[o:o + n] = np.frombuffer(q, dtype=np.uint8)
    return PackedBatch(seq, qual, offsets, lens)


def shard_reads_by_bases(lens, world_size):
    """Contiguous partition of reads into world_size shards balanced by BASES, not reads.

    Returns world_size+1 boundaries b with shard r = reads [b[r], b[r+1]).  Contiguous shards keep the
    reference's output order: rank order == input order (src/seprocessor.cpp:352 deals packs round-robin;
    we deal contiguous blocks and concatenate)."""
    lens = np.asarray(lens, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(lens)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        bounds.append(int(np.searchsorted(csum, target, side="left")))
    bounds.append(len(lens))
    for r in range(1, len(bounds)):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds
