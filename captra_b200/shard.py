"""Multi-GPU plumbing for the tracking batch.

The hot path shards naturally: every op is independent per cloud and trajectories are independent
of each other (model.py:409-478 is batched over them), so each rank tracks a contiguous block of
trajectories with replicated weights and NO data-path collective.  The only exchange is the
end-of-batch sum of the pose-error / loss scalars (what test.py:87-99 accumulates), one
all-reduce of a few floats over NCCL (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous split of `total` trajectories: the first `total % world_size` ranks get one extra."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def category_of(index, num_categories=6):
    """BASELINE cfg4: category = global trajectory index mod 6."""
    return index % num_categories


def pose_error_scalars(pose, gt):
    """Sums (not means) so that the all-reduce of shards equals the single-process value:
    [sum |t - t_gt|, sum |s - s_gt|, sum |R - R_gt|_F^2, count]."""
    t_err = (pose["translation"] - gt["translation"]).norm(dim=-2).sum()
    s_err = (pose["scale"] - gt["scale"]).abs().sum()
    r_err = (pose["rotation"] - gt["rotation"]).pow(2).sum()
    count = torch.tensor(float(pose["scale"].numel()), device=pose["scale"].device)
    return torch.stack([t_err, s_err, r_err, count])


def all_reduce_scalars(vec, group=None):
    """In-place SUM over ranks (no-op when torch.distributed is not initialised)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec
