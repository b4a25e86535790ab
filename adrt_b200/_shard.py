"""Batch sharding helpers for one-process-per-GPU runs.

Batch items never interact in any ADRT routine (reference: `batch` is the
outermost independent loop everywhere, e.g. adrt_cdefs_adrt.hpp:71), so
multi-GPU execution is pure batch sharding with no data-path collective.
``torch.distributed`` is used only for rendezvous, barriers and reducing the
timing / gathering results when the caller wants them in one place.
"""
from __future__ import annotations


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch for `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank {world}/{rank}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """MAX-reduce a scalar (e.g. elapsed milliseconds) over all ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_batch(local, dist=None):
    """All-gather per-rank result shards (tensors with equal trailing dims) along dim 0."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    counts = [int(s.item()) for s in sizes]
    # all_gather needs equal shapes: pad every shard to the largest one
    padded = local.new_zeros((max(counts), *local.shape[1:]))
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:k] for o, k in zip(out, counts)], dim=0)
