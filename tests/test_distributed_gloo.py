"""world_size-2 gloo test of the multi-GPU host logic on CPU: contiguous sharding by bases + the accumulator
all-reduce that replaces Stats::merge / FilterResult::merge.  Each rank runs the oracle on its shard (no GPU here);
the merged blocks must equal a single pass over the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    from fastplong_b200 import distributed as D
    from oracle_lib import OracleEngine
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    batch = cases.ont_batch(3, n=240, mean=1500, p_chimera=0.05)
    mine, bounds = D.shard(batch, rank, world)
    eng = OracleEngine(opt)
    rec = eng.process(mine)
    cyc = D.agree_on_cycles(int(mine.lens.max()) if mine.n_reads else 1)
    blocks = [torch.from_numpy(eng.stats(0, cyc)), torch.from_numpy(eng.stats(1, cyc)), torch.from_numpy(eng.counters())]
    D.allreduce_in_place(blocks)
    allrec = D.gather_records(rec, bounds, rank, world)
    if rank == 0:
        np.savez(tmp, pre=blocks[0].numpy(), post=blocks[1].numpy(), cnt=blocks[2].numpy(), rec=allrec, cyc=cyc,
                 bounds=np.array(bounds))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_merge(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from oracle_lib import OracleEngine, compare_results, compare_stats
    out = str(tmp_path / "merged.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    opt = cases.OPTION_SETS["cut_polyx_cplx"]
    batch = cases.ont_batch(3, n=240, mean=1500, p_chimera=0.05)
    eng = OracleEngine(opt)
    rec = eng.process(batch)
    cyc = int(z["cyc"])
    compare_results(z["rec"], rec, "gathered records")
    compare_stats(z["pre"], eng.stats(0, cyc), "merged pre")
    compare_stats(z["post"], eng.stats(1, cyc), "merged post")
    compare_stats(z["cnt"], eng.counters(), "merged counters")
    b = z["bounds"]
    assert b[0] == 0 and b[-1] == batch.n_reads and 0 < b[1] < batch.n_reads
    # shards are balanced by bases
    left = int(batch.lens[: b[1]].sum())
    assert abs(left - batch.n_bases / 2) < batch.lens.max()


def test_shard_bounds_edge_cases():
    from fastplong_b200.pack import shard_reads_by_bases
    assert shard_reads_by_bases([], 4) == [0, 0, 0, 0, 0]
    assert shard_reads_by_bases([5], 2)[-1] == 1
    b = shard_reads_by_bases([10] * 8, 8)
    assert b == list(range(9))


def test_shard_and_slice_properties():
    """Seeded random length vectors x world sizes: the shards are contiguous, cover every read once, are balanced by bases
    to within one read, and PackedBatch.slice hands back exactly the reads of its range with offsets rebased to 0 (a rank
    uploads its own bytes only) — including zero-length reads and empty shards."""
    import numpy as np
    from fastplong_b200 import pack_reads
    from fastplong_b200.pack import SLOT_ALIGN, shard_reads_by_bases
    rng = np.random.default_rng(17)
    for case in range(40):
        n = int(rng.integers(0, 60))
        lens = np.where(rng.random(n) < 0.15, 0, rng.integers(1, 700, size=n)).astype(np.int64)
        if case % 7 == 0 and n:
            lens[int(rng.integers(n))] = 20000          # one read that outweighs a whole shard
        reads = [(bytes(rng.integers(65, 91, size=int(L), dtype=np.uint8)), bytes(rng.integers(34, 80, size=int(L), dtype=np.uint8))) for L in lens]
        batch = pack_reads(reads) if n else pack_reads([(b"", b"")]).slice(0, 0)
        for world in (1, 2, 3, 8):
            b = shard_reads_by_bases(lens, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            total, biggest = int(lens.sum()), int(lens.max()) if n else 0
            for r in range(world):
                # a boundary sits at the first read whose prefix sum reaches r/world of the bases
                assert abs(int(lens[:b[r]].sum()) - total * r // world) <= biggest
                sh = batch.slice(b[r], b[r + 1])
                assert sh.n_reads == b[r + 1] - b[r]
                if sh.n_reads:
                    assert int(sh.offsets[0]) == 0 and (sh.offsets % SLOT_ALIGN == 0).all()
                    assert int(sh.offsets[-1]) + int(sh.lens[-1]) <= sh.n_bytes
                for i in range(sh.n_reads):
                    assert sh.read(i) == reads[b[r] + i]
