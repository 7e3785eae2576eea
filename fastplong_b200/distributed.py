"""Multi-GPU driver (one process per GPU).  The path shards by read; the only exchange is the final merge
that replaces Stats::merge / FilterResult::merge (src/stats.cpp:1013-1082, src/filterresult.cpp:28-61, called from
src/seprocessor.cpp:108-121).

GPU path: the merge is the C ABI's own (include/fplgpu.h: fpl_comm_init + fpl_allreduce_stats — ncclAllReduce on the
context's stream); torch.distributed is only the out-of-band channel that carries rank 0's 128-byte rendezvous id and
gathers the per-read records.  The same shard / gather logic runs on gloo with host blocks in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist

from .pack import shard_reads_by_bases


def shard(batch, rank, world_size):
    """This rank's contiguous share of the batch, balanced by bases."""
    bounds = shard_reads_by_bases(batch.lens, world_size)
    return batch.slice(bounds[rank], bounds[rank + 1]), bounds


def agree_on_cycles(local_cycles, device=None, group=None):
    """All ranks pad their Stats blocks to the same number of cycles before the all-reduce (host blocks / gloo)."""
    t = torch.tensor([int(local_cycles)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def allreduce_in_place(blocks, group=None):
    """Sum-all-reduce every accumulator block (int64 tensors) in place (host blocks / gloo)."""
    for b in blocks:
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)


def init_engine_comm(engine, rank=None, world_size=None, group=None):
    """Join `engine` (binding.Engine) to the job's NCCL communicator: rank 0 makes the id, everybody receives it
    over the existing torch.distributed group (any backend), every rank calls fpl_comm_init."""
    from .binding import Engine
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    box = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(box[0], rank, world_size)


def merge_engine(engine, cycles=0):
    """Stats::merge / FilterResult::merge over the ranks, in place on the device, on the engine's stream.
    cycles = 0 agrees on the longest read first (one small synchronous all-reduce)."""
    engine.allreduce_stats(cycles)


def gather_records(records, bounds, rank, world_size, group=None):
    """Concatenate per-rank record arrays in rank order == input order (contiguous shards)."""
    out = [None] * world_size
    dist.all_gather_object(out, np.asarray(records), group=group)
    return np.concatenate(out)
