"""Per-frame sharding of a batch over ranks (one process per GPU).

Frames are independent units on this path: the batch index only selects a disjoint slab of the coordinate space
(pcdet/ops/spconv/include/spconv/geometry.h:179-180) and eval-mode BatchNorm has no cross-sample statistics, so
a batch is split by frame with NO collective on the data path.  The only communication is
the reduction of a timing (MAX) or the optional gather of per-rank row counts for reporting.
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world_size):
    """Contiguous frame range [lo, hi) of `rank`; the first n_frames % world_size ranks get one extra frame."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %r/%r" % (rank, world_size))
    base, extra = divmod(int(n_frames), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_frames(frames, rank=None, world_size=None):
    """This rank's frames, re-indexed to local batch ids 0..len-1 (the oracle equivalent is those frames run as
    their own batch)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(len(frames), rank, world_size)
    return frames[lo:hi]


def max_over_ranks(value, device=None):
    """MAX of a python float over all ranks (identity without an initialised process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(counts, device=None):
    """All ranks' row-count lists (for reporting), as a [world, len] list of lists."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(counts)]
    t = torch.tensor(list(counts), dtype=torch.int64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]
