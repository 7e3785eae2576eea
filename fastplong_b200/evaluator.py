"""Adapter auto-detection: the host half of Evaluator::evalAdapterAndReadNum
(src/evaluator.cpp:105-265).  The ten-mer tables come from the device (binding.eval_adapter_kmers ->
fpl_eval_adapter_kmers); what follows them is O(4^10) table work and stays here:

  top_key            Evaluator::getTopKey (src/evaluator.cpp:266-322): the most frequent ten-mer that is not low-complexity
  extend_key         Evaluator::extendKeyToAdapter (:324-407): grow the key base by base, left first, while one
                     neighbour ten-mer clearly dominates and sits one position further
  detect_adapters    the surrounding rules (:150-255): >= 100 reads, count > 10, fold threshold 100, length > 16

Bit-compatible with the reference including its quirks (the "diff" complexity test reads the COUNT's bits, not the
key's; the `starts with GGGG` test compares 8 bits with 0xff).
"""
import numpy as np

KEYLEN = 10
SIZE = 1 << (2 * KEYLEN)
READ_LIMIT = 64 * 1024                 # src/evaluator.cpp:110-111
BASE_LIMIT = 8192 * READ_LIMIT
FOLD_THRESHOLD = 100.0
MAX_LEN = 64
BASES = "ATCG"                         # Evaluator::int2seq (:485-497)


def int2seq(val, n=KEYLEN, is_rna=False):
    bases = "AUCG" if is_rna else BASES
    return "".join(bases[(val >> (2 * (n - 1 - i))) & 3] for i in range(n))


def top_key(counts):
    """Index of the largest count among the ten-mers the reference does not reject; -1 if none (first maximum wins)."""
    k = np.arange(SIZE, dtype=np.int64)
    val = counts.astype(np.int64)
    atcg = np.zeros((4, SIZE), dtype=np.int32)
    for i in range(KEYLEN):
        b = (k >> (2 * i)) & 3
        for c in range(4):
            atcg[c] += (b == c)
    low = (atcg >= KEYLEN - 4).any(axis=0) | ((atcg == 0).sum(axis=0) >= 2)
    low |= (k >> KEYLEN) == (k & ((1 << KEYLEN) - 1))
    # the reference's "diff" loop shifts `val` (the count), not the key: src/evaluator.cpp:296-303
    diff = np.zeros(SIZE, dtype=np.int32)
    for s in range(KEYLEN - 1):
        cur = (val >> ((KEYLEN - s) * 2)) & 3
        last = (val >> ((KEYLEN - s - 1) * 2)) & 3
        diff += (cur != last)
    ok = (diff >= 3) & ~low & (atcg[2] + atcg[3] < KEYLEN - 2) & ((k >> 12) != 0xFF) & (k != 0)
    cand = np.where(ok, val, -1)
    best = int(cand.max())
    if best <= 0:           # `val > topCount` with topCount = 0: a zero count never becomes the top key
        return -1
    return int(np.argmax(cand))


def extend_key(key, counts, position_acc, is_rna=False, left_first=True):
    bases = "AUCG" if is_rna else BASES
    adapter = int2seq(key, KEYLEN, is_rna)
    mask = SIZE - 1
    left_done = right_done = False
    left = bool(left_first)
    while True:
        cur = key
        while len(adapter) < MAX_LEN:
            def nk(b):
                return (b << ((KEYLEN - 1) * 2)) | (cur >> 2) if left else b | (mask & (cur << 2))
            total = sum(int(counts[nk(b)]) for b in range(4))
            extended = False
            for b in range(4):
                n = nk(b)
                cn = int(counts[n])
                if cn == 0:
                    continue
                offset = float(position_acc[n]) / cn - float(position_acc[cur]) / float(counts[cur])
                if cn / float(total) < 0.7:
                    continue
                if cn / float(counts[key]) < 0.5:
                    continue
                if offset > 2 or offset < -4:
                    continue
                cur = n
                extended = True
                adapter = bases[b] + adapter if left else adapter + bases[b]
                break
            if not extended:
                if left:
                    left_done = True
                else:
                    right_done = True
                break
            if len(adapter) == MAX_LEN:
                left_done = right_done = True
                break
        left = not left
        if left_done and right_done:
            break
    return adapter


def evaluated_prefix(lens):
    """How many reads of the input the reference loads: while records < 64 Ki and bases < 512 Mi (:122)."""
    lens = np.asarray(lens, dtype=np.int64)
    csum = np.cumsum(lens)
    over = np.nonzero(csum >= BASE_LIMIT)[0]
    n = int(over[0]) + 1 if len(over) else len(lens)
    return min(n, READ_LIMIT)


def detect_one(counts, position_acc, total, is_rna=False):
    """One side: the adapter string, or None ("Not detected" / too short)."""
    counts = counts.copy()
    total_key = int(np.count_nonzero(counts))
    counts[0] = 0
    key = top_key(counts)
    count = int(counts[key])           # key == -1 reads counts[-1], like the reference's out-of-bounds read would not: guard
    if key < 0:
        return None
    if count > 10 and count * total_key > total * FOLD_THRESHOLD:
        adapter = extend_key(key, counts, position_acc, is_rna, True)
        if len(adapter) > 16:
            return adapter
    return None


def detect_adapters(batch, trim_tail=0, is_rna=False, device=0, kmers=None, pick="abi"):
    """(start, end) as Evaluator::evalAdapterAndReadNum would set opt.adapter.sequenceStart / sequenceEnd when both are
    "auto" ("auto" is kept where nothing is detected).  `batch`: a host PackedBatch holding the head of
    the input; `kmers(batch, side, shift_tail)` supplies the tables (default: the device kernel); pick = "abi": the table
    half is the C ABI's fpl_eval_pick_adapter (C++, host only), "python": detect_one above — the same rules twice, held to
    each other and to the reference binary by tests/test_evaluator.py and tools/fuzz_evaluator_vs_binary.py."""
    from .binding import eval_adapter_kmers, eval_pick_adapter
    n = evaluated_prefix(batch.lens)
    head = batch.slice(0, n)
    if n < 100:
        return "auto", "auto"
    shift_tail = max(1, int(trim_tail))
    fn = kmers or (lambda b, side, st: eval_adapter_kmers(b, side, st, device))
    out = []
    for side in (0, 1):
        counts, acc, total = fn(head, side, shift_tail)
        rna = is_rna if side == 1 else False
        a = eval_pick_adapter(counts, acc, total, rna) if pick == "abi" else detect_one(counts, acc, total, rna)
        out.append(a if a else "auto")
    return out[0], out[1]
