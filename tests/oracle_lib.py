"""TEST INFRASTRUCTURE: ctypes loaders for the CPU checkers.

  OracleEngine  -> oracle/liboracle.so      our plain-C restatement (travels to the GPU box, rebuilt if absent)
  RefEngine     -> oracle/_ref/libfplref.so the unmodified reference's operators (built here from /root/reference;
                                             the prebuilt .so travels to the GPU box)
Both expose the same methods as fastplong_b200.binding.Engine so the parity tests read the same for all three.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from fastplong_b200 import abi
from fastplong_b200.abi import FplAdapters, FplBatch, FplOptions, RESULT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libfplref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "fastplong_ref")


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "fpl_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def have_ref():
    return os.path.exists(REF_SO)


class _CpuEngine:
    prefix = None
    so = None

    def __init__(self, options):
        lib = C.CDLL(self.so)
        p = self.prefix
        self._create = getattr(lib, p + "_create")
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.POINTER(FplOptions), C.POINTER(FplAdapters)]
        self._destroy = getattr(lib, p + "_destroy")
        self._destroy.argtypes = [C.c_void_p]
        self._process = getattr(lib, p + "_process")
        self._process.argtypes = [C.c_void_p, C.POINTER(FplBatch), C.c_void_p]
        self._stats_dl = getattr(lib, p + "_stats_download")
        self._stats_dl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        self._cnt_dl = getattr(lib, p + "_counters_download")
        self._cnt_dl.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        self._segs = getattr(lib, p + "_last_segments")
        self._segs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        self._regs = getattr(lib, p + "_last_mask_regions")
        self._regs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        self.lib = lib
        self.options = options
        o, ad, keep = options.to_abi()
        self._keep = (o, ad, keep)
        self.n_adapters = 2 + len(options.adapter_fasta)
        self.h = self._create(C.byref(o), C.byref(ad))
        self.max_len = 0

    def close(self):
        if self.h:
            self._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, batch):
        res = np.zeros(batch.n_reads, dtype=RESULT_DTYPE)
        b = batch.to_abi()
        rc = self._process(self.h, C.byref(b), res.ctypes.data)
        assert rc == 0
        if batch.n_reads:
            self.max_len = max(self.max_len, int(batch.lens.max()))
        return res

    def _list(self, fn, dtype):
        n = C.c_int64()
        fn(self.h, None, 0, C.byref(n))
        out = np.zeros(n.value, dtype=dtype)
        if n.value:
            assert fn(self.h, out.ctypes.data, n.value, C.byref(n)) == 0
        return out

    def segments(self):
        """--mask/--break: every output read of the last process() call."""
        return self._list(self._segs, abi.SEGMENT_DTYPE)

    def mask_regions(self):
        return self._list(self._regs, abi.REGION_DTYPE)

    def stats(self, which, cycles=None):
        cyc = int(cycles if cycles is not None else max(self.max_len, 1))
        out = np.zeros(abi.stats_words(cyc), dtype=np.int64)
        rc = self._stats_dl(self.h, which, out.ctypes.data, cyc)
        assert rc == 0, rc
        return out

    def counters(self):
        n = abi.counter_words(self.n_adapters)
        out = np.zeros(n, dtype=np.int64)
        rc = self._cnt_dl(self.h, out.ctypes.data, n)
        assert rc == 0, rc
        return out


class OracleEngine(_CpuEngine):
    prefix = "orc"

    def __init__(self, options):
        self.so = build_oracle()
        super().__init__(options)
        self.lib.orc_edit_distance.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        self.lib.orc_search_adapter.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.orc_trim_and_cut.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.lib.orc_trim_polyx.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]

    def edit_distance(self, a, b):
        return self.lib.orc_edit_distance(a, len(a), b, len(b))

    def search_adapter(self, read, adapter, start=0, length=-1, left=False, right=False):
        return self.lib.orc_search_adapter(self.h, read, len(read), adapter, len(adapter), start, length,
                                           int(left), int(right))

    def trim_and_cut(self, seq, qual):
        lo, n = C.c_int(), C.c_int()
        dropped = self.lib.orc_trim_and_cut(self.h, seq, qual, len(seq), C.byref(lo), C.byref(n))
        return None if dropped else (lo.value, n.value)

    def trim_polyx(self, seq, min_len):
        base, plen = C.c_int(), C.c_int()
        nl = self.lib.orc_trim_polyx(seq, len(seq), min_len, C.byref(base), C.byref(plen))
        return nl, base.value, plen.value


class RefEngine(_CpuEngine):
    prefix = "ref"
    so = REF_SO

    def __init__(self, options):
        super().__init__(options)
        self.lib.ref_edit_distance.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        self.lib.ref_search_adapter.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_double, C.c_int, C.c_int,
                                                C.c_int, C.c_int]

    def edit_distance(self, a, b):
        return self.lib.ref_edit_distance(a, len(a), b, len(b))

    def search_adapter(self, read, adapter, start=0, length=-1, left=False, right=False):
        return self.lib.ref_search_adapter(read, len(read), adapter, self.options.distance_threshold, start, length,
                                           int(left), int(right))


def compare_results(a, b, what="results"):
    """Field-by-field equality of two RESULT_DTYPE arrays with a readable failure."""
    assert a.shape == b.shape
    for name in RESULT_DTYPE.names:
        x, y = a[name], b[name]
        if not np.array_equal(x, y):
            bad = np.nonzero((x != y).reshape(len(a), -1).any(axis=1))[0]
            i = int(bad[0])
            raise AssertionError(f"{what}: field {name} differs on {len(bad)} reads, first read {i}: "
                                 f"{a[i]} vs {b[i]}")


def compare_lists(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    for name in a.dtype.names:
        if not np.array_equal(a[name], b[name]):
            i = int(np.nonzero(a[name] != b[name])[0][0])
            raise AssertionError(f"{what}: field {name} differs first at {i}: {a[i]} vs {b[i]}")


def compare_stats(a, b, what="stats"):
    assert a.shape == b.shape, (a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.nonzero(a != b)[0]
        raise AssertionError(f"{what}: {len(bad)} words differ, first at {int(bad[0])}: "
                             f"{int(a[bad[0]])} vs {int(b[bad[0]])}")


def kmer10_tables(batch, side, shift_tail=1):
    """TEST INFRASTRUCTURE: the ten-mer tables of Evaluator::evalAdapterAndReadNum (src/evaluator.cpp:176-191,219-234)
    in plain numpy — the checker of fpl_eval_adapter_kmers and the table source of the CPU tests of evaluator.py."""
    code = np.full(256, -1, dtype=np.int64)
    for ch, v in ((b"A", 0), (b"T", 1), (b"U", 1), (b"C", 2), (b"G", 3)):
        code[ch[0]] = v
    counts = np.zeros(1 << 20, dtype=np.uint32)
    acc = np.zeros(1 << 20, dtype=np.uint64)
    total = 0
    for i in range(batch.n_reads):
        o, n = int(batch.offsets[i]), int(batch.lens[i])
        last = n - 10 - shift_tail
        p0, p1 = (0, min(last, 127)) if side == 0 else (max(0, last - 128), last)
        if p1 < p0:
            continue
        c = code[batch.seq[o + p0:o + p1 + 10]]
        win = np.lib.stride_tricks.sliding_window_view(c, 10)
        ok = (win >= 0).all(axis=1)
        keys = (win * (4 ** np.arange(9, -1, -1))).sum(axis=1)[ok]
        pos = np.arange(p0, p1 + 1)[ok]
        np.add.at(counts, keys, 1)
        np.add.at(acc, keys, (pos if side == 0 else n - pos).astype(np.uint64))
        total += int(ok.sum())
    return counts, acc, total
