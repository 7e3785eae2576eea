"""Multi-GPU driver logic (one process per GPU, torch.distributed): the path shards by read, the only exchange is
the final merge that replaces Stats::merge / FilterResult::merge (src/stats.cpp:1013-1082, src/filterresult.cpp:28-61).

Works with any backend: NCCL on the device blocks exposed by the C ABI (fpl_stats_device_ptr), or gloo on host copies
(the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

from .pack import shard_reads_by_bases


def shard(batch, rank, world_size):
    """This rank's contiguous share of the batch, balanced by bases."""
    bounds = shard_reads_by_bases(batch.lens, world_size)
    return batch.slice(bounds[rank], bounds[rank + 1]), bounds


def agree_on_cycles(local_cycles, device=None, group=None):
    """All ranks pad their Stats blocks to the same number of cycles before the all-reduce."""
    t = torch.tensor([int(local_cycles)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def allreduce_in_place(blocks, group=None):
    """Sum-all-reduce every accumulator block (int64 tensors) in place."""
    for b in blocks:
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)


def merge_engine(engine, device, group=None):
    """NCCL path: all-reduce the engine's device-resident accumulators in place (pre, post, counters)."""
    cyc = agree_on_cycles(engine.cycles, device=device, group=group)
    engine.reserve_cycles(cyc)
    engine.sync()
    blocks = [torch.as_tensor(engine.stats_device(0), device=device), torch.as_tensor(engine.stats_device(1), device=device),
              torch.as_tensor(engine.counters_device(), device=device)]
    allreduce_in_place(blocks, group)
    return blocks


def gather_records(records, bounds, rank, world_size, group=None):
    """Concatenate per-rank record arrays in rank order == input order (contiguous shards)."""
    out = [None] * world_size
    dist.all_gather_object(out, np.asarray(records), group=group)
    return np.concatenate(out)
