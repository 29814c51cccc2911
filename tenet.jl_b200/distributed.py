"""Multi-GPU execution: one process per GPU, slices dealt round-robin, ONE all-reduce at the end.

Slice id s goes to rank s mod world_size; every rank holds all leaves (they are tiny next to the
intermediates), contracts its slices into a local accumulator and the accumulators are summed with a single
NCCL all-reduce over NVLink (libtnb200's tnb_comm_allreduce_sum; the 128-byte NCCL id travels through
torch.distributed's store, which is plumbing only).  Un-sliced networks do not shard: replicas only.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def slice_range_for_rank(nslices: int, rank: int, world: int, limit: Optional[int] = None):
    """(begin, step, end) of the round-robin share of `rank`; `limit` caps slices per rank (bench sampling)."""
    end = nslices
    if limit is not None:
        end = min(nslices, rank + world * limit)
    return rank, world, end


def slices_of_rank(nslices: int, rank: int, world: int, limit: Optional[int] = None):
    b, s, e = slice_range_for_rank(nslices, rank, world, limit)
    return list(range(b, e, s))


def init_comm(ctx, rank: Optional[int] = None, world: Optional[int] = None):
    """Create the NCCL communicator of `ctx`; the unique id is broadcast with torch.distributed (any backend)."""
    import torch.distributed as dist
    if rank is None or world is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)


def allreduce_sum(ctx, array, world: Optional[int] = None):
    """In-place sum of a dense B200Array over all ranks (stream-ordered).  `world`: the rank count the caller's
    launcher reports; a context whose communicator has a different size (init_comm never ran on it, or ran on
    another context) would silently hand back a PARTIAL sum, so that is an error."""
    from ._lib import check, TnbError, TNB_ENCCL
    if world is not None and world > 1:
        have = int(ctx.lib.tnb_comm_size(ctx.handle))
        if have != world:
            raise TnbError(TNB_ENCCL, f"all-reduce over {world} ranks requested but this context's communicator has "
                                      f"{have} rank(s): call distributed.init_comm(ctx) first")
    check(ctx.handle, ctx.lib.tnb_comm_allreduce_sum(ctx.handle, array.buffer.handle,
                                                     array.offset * array.dtype.itemsize, array.size, array.dtype_code))


def contract_distributed(plan, ctx, limit: Optional[int] = None, reduce: bool = True):
    """Run this rank's share of the plan's slices and all-reduce the accumulator."""
    rank, world = rank_world()
    b, s, e = slice_range_for_rank(plan.nslices, rank, world, limit)
    plan.zero_output()
    plan.execute(b, s, e, accumulate=True)
    if reduce and world > 1:
        allreduce_sum(ctx, plan.out_array, world)
    return plan.result()


def host_allreduce_sum(x: np.ndarray) -> np.ndarray:
    """gloo/CPU path used by the world_size-2 tests: sums host partial results with torch.distributed."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    if np.iscomplexobj(x):
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.float64 if x.dtype == np.complex128 else np.float32).copy())
        dist.all_reduce(t)
        return t.numpy().view(x.dtype).reshape(x.shape)
    t = torch.from_numpy(np.ascontiguousarray(x).copy())
    dist.all_reduce(t)
    return t.numpy()
