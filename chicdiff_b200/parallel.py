"""Sharded runs: one process per GPU, regions partitioned by bait.

The reference has no distributed mode (its `parallel=TRUE` only means "read each CHiCAGO file once",
chicdiff.R:1464-1466).  Here the region universe is cut into contiguous bait-aligned shards
(cd_plan_shards); each rank holds every row of its baits for every sample, so aggregation needs no
exchange, and the global steps inside cd_region_test (size-factor medians, dispersion trend + MAD,
theta-grid deviances) use the NCCL communicator created below.  torch.distributed is only the plumbing
that ships the 128-byte NCCL id and collects the per-shard output tables.
"""
import numpy as np


def shard_slices(region_bait, row_off, world):
    """Region bounds of every shard: bounds[k] .. bounds[k+1] (bait aligned, balanced by rows)."""
    from . import engine
    return engine.plan_shards(region_bait, row_off, world)


def take_shard(row_off, cols, bounds, rank):
    """Local CSR offsets (re-based to 0) and the row slices of per-row column arrays (last axis = rows)."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    r0, r1 = int(row_off[lo]), int(row_off[hi])
    local_off = np.ascontiguousarray(row_off[lo:hi + 1] - r0)
    local_cols = [np.ascontiguousarray(c[..., r0:r1]) for c in cols]
    return local_off, local_cols, (lo, hi)


def init_comm(eng, dist=None):
    """Joins this rank's context to the NCCL communicator; torch.distributed broadcasts the id."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1, 0
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.comm_init(world, rank, box[0])
    return world, rank


GLOBAL_KEYS = ("sizeFactors", "deviances", "theta", "trend_a0", "trend_a1", "varLogDispEsts", "dispPriorVar")


def gather_columns(local, dist=None, dst=0, global_keys=GLOBAL_KEYS):
    """Concatenates per-region result columns of all shards in rank order on `dst` (None elsewhere).

    local: dict name -> array whose last axis is the shard's regions; entries named in `global_keys`
    are results of the global steps (identical on every rank) and are passed through."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [None] * world
    dist.all_gather_object(box, local)      # (NCCL process groups have no gather_object)
    if rank != dst:
        return None
    out = {}
    for k in local:
        if k not in global_keys and isinstance(local[k], np.ndarray) and local[k].ndim >= 1:
            out[k] = np.concatenate([b[k] for b in box], axis=-1)
        else:
            out[k] = box[0][k]          # scalars of the global steps are identical on every rank
    return out
