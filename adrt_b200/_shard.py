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


# ---------------------------------------------------------------------------
# One large image over several GPUs: quadrant sharding
# ---------------------------------------------------------------------------
# The four quadrants of adrt are independent transforms of four orientations of
# the image (adrt_cdefs_adrt.hpp:124-175) and bdrt treats planes independently,
# so `truncate(bdrt(adrt(x)))` splits 4 ways with the image replicated and ONE
# exchange per application: an all-gather of the four truncated (n, n)
# back-projections, summed locally in the fixed order ((t0+t1)+t2)+t3 so that
# the result is bit-identical to the single-GPU `np.mean(..., axis=-3)` order.

def quadrant_owner_range(world: int, rank: int) -> tuple[int, int]:
    """Quadrants [q_first, q_first + q_count) owned by `rank` when `world` in {1, 2, 4}
    ranks share one image (ranks >= 4 own nothing: angle-block sharding inside a
    quadrant is not implemented yet)."""
    if world >= 4:
        return (rank, 1) if rank < 4 else (0, 0)
    per = 4 // world
    return rank * per, per


def truncate_quadrant(z, q: int):
    """utils.truncate for a single quadrant plane ``(..., 2n-1, n)`` -> ``(..., n, n)``."""
    n = z.shape[-1]
    sq = z[..., :n, :]
    if q == 0:
        return sq.flip(-2).transpose(-1, -2)
    if q == 1:
        return sq.flip(-2)
    if q == 2:
        return sq
    return sq.flip((-1, -2)).transpose(-1, -2)


def _local_quadrant_backprojections(x, q_first, q_count):
    from . import _adrt_cdefs as cd

    y = cd.adrt_quadrants(x, q_first, q_count)
    z = cd.bdrt_planes(y, rows=x.shape[-1])
    return [truncate_quadrant(z[..., i, :, :], q_first + i).contiguous() for i in range(q_count)]


def sharded_normal_operator(x, dist=None, *, local_fn=_local_quadrant_backprojections):
    """``mean_q(truncate(bdrt(adrt(x))))`` with the quadrants spread over the ranks of
    the default process group.  `x` ``(n, n)`` or ``(B, n, n)`` must be replicated
    on every rank; the result is replicated too and bit-identical to the
    single-GPU ``recipes.normal_operator``.  `local_fn(x, q_first, q_count)` returns
    the rank's truncated back-projections (overridable so that CPU tests can
    exercise the exchange with the oracle)."""
    import torch

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    q_first, q_count = quadrant_owner_range(world, rank)
    mine = local_fn(x, q_first, q_count) if q_count else []
    if world == 1:
        parts = mine
    else:
        per = max(1, 4 // min(world, 4))
        # fixed-size buffer per rank: `per` images, zero for ranks that own nothing
        local = torch.zeros((per, *x.shape), dtype=x.dtype, device=x.device)
        for i, t in enumerate(mine):
            local[i] = t
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        parts = []
        for r in range(min(world, 4)):
            qf, qc = quadrant_owner_range(world, r)
            parts += [gathered[r][i] for i in range(qc)]
    t0, t1, t2, t3 = parts
    return (((t0 + t1) + t2) + t3) / 4


def bind_host_to_device(index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to CUDA device
    `index` (its NUMA node), so that pinned host buffers allocated afterwards and
    the staging threads of the host API sit next to the GPU's PCIe root.  One
    process per GPU should call this before allocating host memory.  Returns
    False (and changes nothing) when NVML or the affinity call is unavailable."""
    import os

    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(index)
        bus_id = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
