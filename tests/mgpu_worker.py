"""Worker of tests/test_gpu_multigpu.py, started by `python -m torch.distributed.run --nproc-per-node N`: shard ONE
seeded batch by bases, process the shard on this rank's GPU through the C ABI, merge the accumulators with
fpl_allreduce_stats (NCCL), gather the records, and on rank 0 compare everything with a single pass of the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import cases
    from fastplong_b200 import distributed as D
    from fastplong_b200.binding import Engine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")           # out-of-band channel only; the data path uses the C ABI's communicator
    opt = cases.OPTION_SETS[sys.argv[1] if len(sys.argv) > 1 else "cut_polyx_cplx"]
    opt.device = local
    batch = cases.ont_batch(11, n=600, mean=2500, p_chimera=0.05)
    mine, bounds = D.shard(batch, rank, world)
    eng = Engine(opt)
    D.init_engine_comm(eng, rank, world)
    rec = eng.process(mine)
    D.merge_engine(eng)                        # agree on the cycle count, then one NCCL group on the engine's stream
    eng.sync()
    cyc = int(batch.lens.max())
    pre, post, cnt = eng.stats(0, cyc), eng.stats(1, cyc), eng.counters()
    allrec = D.gather_records(rec, bounds, rank, world)
    # a second merge with an agreed cycle count and nothing new accumulated multiplies every word by the world size
    eng.allreduce_stats(cyc)
    eng.sync()
    again = eng.stats(0, cyc)
    ok = True
    if rank == 0:
        from oracle_lib import OracleEngine, compare_results, compare_stats
        orc = OracleEngine(opt)
        ref = orc.process(batch)
        compare_results(allrec, ref, "gathered records")
        compare_stats(pre, orc.stats(0, cyc), "merged pre")
        compare_stats(post, orc.stats(1, cyc), "merged post")
        compare_stats(cnt, orc.counters(), "merged counters")
        assert np.array_equal(again, pre * world), "second all-reduce"
        assert 0 < bounds[1] < batch.n_reads
        print("MGPU_OK world=%d reads=%d" % (world, batch.n_reads))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
