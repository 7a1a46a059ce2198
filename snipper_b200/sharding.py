"""Snippet sharding for multi-GPU runs (one process per GPU).

Inference partitions independent snippets across ranks with NO data-path collective; the only
communication is the timing reduction (max over ranks).  Training wraps the model in stock DDP.
Works on any backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of ``n_items`` for ``rank`` (first ranks get the extra)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(items_this_rank, elapsed_s_this_rank, device="cpu"):
    """Whole-job throughput = all items processed / slowest rank's time."""
    total = sum_over_ranks(items_this_rank, device)
    slowest = max_over_ranks(elapsed_s_this_rank, device)
    return total / slowest if slowest > 0 else float("inf")
