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
# One large image over several GPUs (BASELINE config 5)
# ---------------------------------------------------------------------------
# Two levels.  (i) The four quadrants of adrt are independent transforms of four
# orientations of the image (adrt_cdefs_adrt.hpp:124-175) and bdrt treats planes
# independently, so up to 4 ranks each take whole quadrants.  (ii) Beyond 4 ranks
# every quadrant is shared by `parts` = world / 4 ranks by ANGLE BLOCK: a butterfly
# stage only pairs rows of adjacent blocks with the same incoming angle
# (adrt_cdefs_adrt.hpp:78-89), so a rank runs all passes but the last on its own
# image-row blocks, the ranks of the quadrant swap workspace rows ONCE (row (blk, a)
# goes from the owner of block blk to the owner of angle a), and the last pass runs
# on the rank's own angles; bdrt mirrors this (include/adrt_b200.h, adrt_b200_adrt_part).
# `truncate(bdrt(adrt(x)))` then needs one all-gather of the ranks' (n, n / parts)
# pieces of the truncated back-projections; every rank sums the four quadrants in
# the fixed order ((t0+t1)+t2)+t3 so the result is bit-identical to the single-GPU
# `np.mean(..., axis=-3)` order.

SUPPORTED_WORLDS = (1, 2, 4, 8, 16, 32)
M_LAST = 3   # stages fused by the pass after the exchange: 8 adjacent sinogram columns per group (one 32-byte sector in fp32)


def image_layout(world: int, parts: int | None = None) -> tuple[int, int]:
    """(quadrants per rank group, ranks per group) when `world` ranks share one image.

    The ranks form ``world / parts`` groups of `parts` consecutive ranks; a group works on
    ``4 / groups`` quadrants and splits each of them into `parts` angle blocks.  Default:
    whole quadrants up to 4 ranks (parts = 1), 4 groups beyond (8 ranks: 4 quadrants x 2
    angle halves).  `parts` (or $ADRT_B200_SHARD_PARTS) overrides, e.g. 2 ranks with
    parts = 2: one group, every quadrant split in two."""
    import os

    if world not in SUPPORTED_WORLDS:
        raise ValueError(f"single-image sharding supports world sizes {SUPPORTED_WORLDS}, got {world}")
    if parts is None and os.environ.get("ADRT_B200_SHARD_PARTS"):
        parts = int(os.environ["ADRT_B200_SHARD_PARTS"])
    if parts is None:
        parts = 1 if world <= 4 else world // 4
    groups = world // parts if parts >= 1 and world % parts == 0 else 0
    if groups not in (1, 2, 4) or parts > (1 << M_LAST):
        raise ValueError(f"cannot split {world} ranks into groups of {parts}: need 1, 2 or 4 groups of at most {1 << M_LAST} ranks")
    return 4 // groups, parts


def quadrant_owner_range(world: int, rank: int, parts: int | None = None) -> tuple[int, int]:
    """Quadrants [q_first, q_first + q_count) that `rank` works on."""
    per, parts = image_layout(world, parts)
    if not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} for world {world}")
    return (rank // parts) * per, per


def part_of(world: int, rank: int, parts: int | None = None) -> tuple[int, int]:
    """(part, parts): the rank's angle block inside its quadrants."""
    _, parts = image_layout(world, parts)
    return rank % parts, parts


def orient_square(sq, q: int):
    """Undo adrt_init's orientation on the top ``(..., n, n)`` square of quadrant `q`
    (the per-quadrant piece of utils.truncate, utils.py:234-242)."""
    if q == 0:
        return sq.flip(-2).transpose(-1, -2)
    if q == 1:
        return sq.flip(-2)
    if q == 2:
        return sq
    return sq.flip((-1, -2)).transpose(-1, -2)


def truncate_quadrant(z, q: int):
    """utils.truncate for a single quadrant plane ``(..., 2n-1, n)`` -> ``(..., n, n)``."""
    n = z.shape[-1]
    return orient_square(z[..., :n, :], q)


def _local_rows(x, q_first, q_count, part, parts, base_rank, dist):
    """The rank's share of ``bdrt(adrt(x))``: offsets ``d < n``, quadrants
    ``q_first .. q_first+q_count-1``, columns ``[part*n/parts, (part+1)*n/parts)``:
    ``(B, q_count, n, n/parts)`` for `x` ``(B, n, n)``."""
    from . import _adrt_cdefs as cd

    n = x.shape[-1]
    if parts > 1:
        return _part_backprojection(x, q_first, q_count, part, parts, base_rank, dist)
    if n < 2:
        z = cd.bdrt_planes(cd.adrt_quadrants(x, q_first, q_count), rows=n)
    else:
        z = cd.adrt_bdrt_rows(x, q_first, q_count)   # sinogram handed over as workspace rows
    return z[:, :, :n, :].contiguous()


def _finish_mean(zfull):
    """``mean_q(truncate(z))`` with the single-GPU kernel (NumPy's summation order)."""
    from . import _adrt_cdefs as cd

    return cd.truncate_mean(zfull, 1.0)


def exchange_rows(xbuf, part: int, parts: int, base_rank: int, dist, forward: bool, cols=None) -> None:
    """The one data-path exchange of angle-block sharding, in place on `xbuf`
    ``(planes, blocks, angles, pitch)``.  Rank ``base_rank + p`` owns blocks
    ``[p*blocks/parts, (p+1)*blocks/parts)`` and angles ``[p*angles/parts, ...)``.
    forward (adrt): a rank holds all angles of its blocks and needs all blocks of its
    angles; transposed (bdrt): the other way round.  One batched send/recv per peer.
    `cols`: only the first `cols` elements of every row travel (row-limited bdrt)."""
    import torch

    nblk, e = xbuf.shape[1], xbuf.shape[2]
    bl = lambda p: slice(p * nblk // parts, (p + 1) * nblk // parts)   # noqa: E731
    an = lambda p: slice(p * e // parts, (p + 1) * e // parts)         # noqa: E731
    if cols is not None and cols < xbuf.shape[3]:
        xbuf = xbuf[..., :cols]
    ops, recvs = [], []
    for p in range(parts):
        if p == part:
            continue
        if forward:
            send = xbuf[:, bl(part), an(p)].contiguous()
            target = xbuf[:, bl(p), an(part)]
        else:
            send = xbuf[:, bl(p), an(part)].contiguous()
            target = xbuf[:, bl(part), an(p)]
        recv = torch.empty(target.shape, dtype=xbuf.dtype, device=xbuf.device)
        ops.append(dist.P2POp(dist.isend, send, base_rank + p))
        ops.append(dist.P2POp(dist.irecv, recv, base_rank + p))
        recvs.append((target, recv))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for target, recv in recvs:
        target.copy_(recv)


def _part_backprojection(x, q_first: int, q_count: int, part: int, parts: int, base_rank: int, dist, exchange=None):
    """Offsets ``d < n``, columns ``[part*n/parts, (part+1)*n/parts)`` of quadrants
    ``q_first .. q_first+q_count-1`` of ``bdrt(adrt(x))`` for `x` ``(B, n, n)``:
    ``(B, q_count, n, n/parts)``, computed by the `parts` ranks ``base_rank ..`` together
    (two exchanges: one inside adrt, one inside bdrt)."""
    exchange = exchange or exchange_rows
    import torch

    from . import _lib

    lib = _lib.load()
    B, n = int(x.shape[0]), int(x.shape[-1])
    code = _lib.F32 if x.dtype == torch.float32 else _lib.F64
    m_last = M_LAST
    pf = int(lib.adrt_b200_part_exchange_pitch(n, code, m_last, 1))
    pb = int(lib.adrt_b200_part_exchange_pitch(n, code, m_last, 0))
    if pf == 0 or pb == 0 or (1 << m_last) < parts:
        raise ValueError(f"angle-block sharding needs n >= {1 << (m_last + 1)} and at most {1 << m_last} ranks per quadrant")
    nblk, e = 1 << m_last, n >> m_last
    D = 2 * n - 1
    dev = x.device
    nbytes = int(lib.adrt_b200_part_workspace_bytes(B * q_count, n, code, m_last))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        planes = B * q_count
        xf = torch.empty((planes, nblk, e, pf), dtype=x.dtype, device=dev)
        sino = torch.empty((planes, D, n), dtype=x.dtype, device=dev)
        args = (B, n, code, q_first, q_count, part, parts, m_last)
        _lib.check(lib.adrt_b200_adrt_part(x.data_ptr(), xf.data_ptr(), None, *args, 0, ws.data_ptr(), nbytes, stream), "adrt_part")
        exchange(xf, part, parts, base_rank, dist, True)
        _lib.check(lib.adrt_b200_adrt_part(None, xf.data_ptr(), sino.data_ptr(), *args, 1, ws.data_ptr(), nbytes, stream), "adrt_part")
        del xf
        xb = torch.empty((planes, nblk, e, pb), dtype=x.dtype, device=dev)
        out = torch.empty((planes, D, n), dtype=x.dtype, device=dev)
        bargs = (planes, n, n, code, part, parts, m_last)
        _lib.check(lib.adrt_b200_bdrt_part(sino.data_ptr(), xb.data_ptr(), None, *bargs, 0, ws.data_ptr(), nbytes, stream), "bdrt_part")
        exchange(xb, part, parts, base_rank, dist, False, int(lib.adrt_b200_part_exchange_cols(n, code, m_last, n)))
        _lib.check(lib.adrt_b200_bdrt_part(None, xb.data_ptr(), out.data_ptr(), *bargs, 1, ws.data_ptr(), nbytes, stream), "bdrt_part")
    cols = slice(part * n // parts, (part + 1) * n // parts)
    return out.view(B, q_count, D, n)[:, :, :n, cols].contiguous()


def _slab_mean(piece, rank: int, world: int, per: int, parts: int, dist):
    """Quadrant mean of the ranks' shares WITHOUT gathering them everywhere.

    `piece` ``(B, per, n, w)``: this rank's offsets ``d < n`` / columns ``[p*w, (p+1)*w)`` of its
    group's quadrants.  The result is cut into `world` slabs of ``h = n / world`` image rows; rank k
    receives exactly the four pieces of the truncated quadrants that cover its slab -- row blocks of
    the untransposed quadrants 1, 2 from the `parts` ranks of those quadrants, one column block of
    each transposed quadrant 0, 3 (utils.truncate orientation, utils.py:234-242) -- sums them in
    NumPy's order ``((t0 + t1) + t2) + t3`` and divides by 4 (the arithmetic of the ``truncate_mean``
    kernel with divisor 1, hence the same bits), and ONE all-gather of the slabs replicates the
    result.  Per rank ``4 n^2 / world`` elements arrive instead of ``4 n^2 (world - 1) / world``."""
    import torch

    B, n, w = int(piece.shape[0]), int(piece.shape[2]), int(piece.shape[3])
    h = n // world
    grp_of = lambda q: q // per                 # noqa: E731
    my_grp, my_p = divmod(rank, parts)
    my_quads = range(my_grp * per, (my_grp + 1) * per)

    def block(q, k):
        """What the owner of (quadrant q, column block my_p) holds for slab k, or None."""
        i = q - my_grp * per
        if q == 2:
            return piece[:, i, k * h:(k + 1) * h, :]
        if q == 1:
            return piece[:, i, n - (k + 1) * h:n - k * h, :]
        lo = k * h if q == 0 else n - (k + 1) * h           # the columns of z_q that slab k needs
        if lo // w != my_p:
            return None
        return piece[:, i, :, lo - my_p * w:lo - my_p * w + h]

    def block_shape(q, k, p):
        if q in (1, 2):
            return (B, h, w)
        lo = k * h if q == 0 else n - (k + 1) * h
        return (B, n, h) if lo // w == p else None

    # messages: one flat tensor per (source, destination) pair, the source's blocks for that slab in quadrant order
    sends, recvs, ops = {}, {}, []
    for k in range(world):
        blocks = [b for b in (block(q, k) for q in my_quads) if b is not None]
        if blocks:
            sends[k] = torch.cat([b.reshape(-1) for b in blocks])
    for src in range(world):
        g, p = divmod(src, parts)
        shapes = [(q, sh) for q in range(g * per, (g + 1) * per) for sh in [block_shape(q, rank, p)] if sh is not None]
        if shapes:
            recvs[src] = shapes
    bufs = {}
    for src, shapes in recvs.items():
        if src == rank:
            bufs[src] = sends[rank]
            continue
        count = sum(a * b * c for _, (a, b, c) in shapes)
        bufs[src] = torch.empty(count, dtype=piece.dtype, device=piece.device)
        ops.append(dist.P2POp(dist.irecv, bufs[src], src))
    for k, t in sends.items():
        if k != rank:
            ops.append(dist.P2POp(dist.isend, t, k))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    # assemble the four oriented slabs (B, h, n)
    T = [None] * 4
    for src, shapes in recvs.items():
        _, p = divmod(src, parts)
        off = 0
        for q, sh in shapes:
            cnt = sh[0] * sh[1] * sh[2]
            raw = bufs[src][off:off + cnt].view(sh)
            off += cnt
            if q in (1, 2):
                if T[q] is None:
                    T[q] = torch.empty((B, h, n), dtype=piece.dtype, device=piece.device)
                T[q][:, :, p * w:(p + 1) * w] = raw.flip(-2) if q == 1 else raw
            elif q == 0:
                T[0] = raw.flip(-2).transpose(-1, -2)
            else:
                T[3] = raw.flip((-1, -2)).transpose(-1, -2)
    slab = ((((T[0] + T[1]) + T[2]) + T[3]) / 4).contiguous()
    out = torch.empty((world, B, h, n), dtype=piece.dtype, device=piece.device)
    try:
        dist.all_gather_into_tensor(out, slab)
    except (RuntimeError, NotImplementedError):
        lst = [torch.empty_like(slab) for _ in range(world)]
        dist.all_gather(lst, slab)
        out = torch.stack(lst, dim=0)
    return out.permute(1, 0, 2, 3).reshape(B, n, n)


def sharded_normal_operator(x, dist=None, *, local_fn=_local_rows, finish_fn=_finish_mean, parts=None):
    """``mean_q(truncate(bdrt(adrt(x))))`` with ONE image (or batch) spread over the ranks
    of the default process group: by quadrant up to 4 ranks, by quadrant x angle block
    beyond (8 ranks = 4 quadrants x 2 angle halves).  `x` ``(n, n)`` or ``(B, n, n)`` must
    be replicated on every rank; the result is replicated too and bit-identical to the
    single-GPU ``recipes.normal_operator``: every rank back-projects its share, receives
    the pieces of the four truncated quadrants that cover its slab of the result, sums them in
    the ``truncate_mean`` kernel's order, and one all-gather of the slabs replicates the result
    (``_slab_mean``, the default from 4 ranks on; with 2 ranks, or with ``ADRT_B200_SHARD_GATHER=1``, the
    older form: all-gather of all shares + the ``truncate_mean`` kernel over them;
    ``ADRT_B200_SHARD_GATHER=0`` forces the slab form).
    `local_fn` / `finish_fn` are overridable so that CPU tests can exercise the exchanges
    with the oracle."""
    import torch

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    per, parts = image_layout(world, parts)
    q_first, q_count = quadrant_owner_range(world, rank, parts)
    part = rank % parts
    squeeze = x.ndim == 2
    xb = (x[None] if squeeze else x).contiguous()
    B, n = int(xb.shape[0]), int(xb.shape[-1])
    w = n // parts
    piece = local_fn(xb, q_first, q_count, part, parts, rank - part, dist)       # (B, per, n, w)
    import os

    # measured on one 8192^2 fp32 image (profiles/r03_normal_op_sharded_8192.jsonl, gpurun_out/r2d_*): the slab
    # exchange wins from 4 ranks on (4: 2.59 vs 2.95 ms, 8: 2.95 vs 3.69 ms); with 2 ranks its assembly passes over
    # half-image slabs cost more than the smaller transfer saves (4.70 vs 4.03 ms)
    gather = os.environ.get("ADRT_B200_SHARD_GATHER")
    use_slab = (world >= 4) if gather is None else gather == "0"
    if world > 1 and finish_fn is _finish_mean and n % world == 0 and n // world >= 1 and use_slab:
        # every rank receives only what covers its slab of the result (_slab_mean)
        res = _slab_mean(piece, rank, world, per, parts, dist)
        return res[0] if squeeze else res
    if world == 1:
        flat = piece[None]
    else:
        flat = torch.empty((world, *piece.shape), dtype=piece.dtype, device=piece.device)
        try:
            dist.all_gather_into_tensor(flat, piece.contiguous())
        except (RuntimeError, NotImplementedError):   # backends without the flat form
            parts_list = [torch.empty_like(piece) for _ in range(world)]
            dist.all_gather(parts_list, piece.contiguous())
            flat = torch.stack(parts_list, dim=0)
    if finish_fn is _finish_mean and flat.is_cuda and w >= 32 and flat.dtype in (torch.float32, torch.float64):
        # the quadrant mean reads the gathered shares where they lie (no reassembly pass)
        from . import _lib

        lib = _lib.load()
        res = torch.empty((B, n, n), dtype=flat.dtype, device=flat.device)
        with torch.cuda.device(flat.device):
            rc = lib.adrt_b200_truncate_mean_shares(flat.data_ptr(), res.data_ptr(), B, n, per, parts, 1.0,
                                                    _lib.F32 if flat.dtype == torch.float32 else _lib.F64,
                                                    torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "truncate_mean_shares")
        return res[0] if squeeze else res
    # the rows truncate keeps, in the public sinogram layout (rows d >= n are never read)
    zfull = torch.empty((B, 4, 2 * n - 1, n), dtype=piece.dtype, device=piece.device)
    for r in range(world):
        grp, p = divmod(r, parts)
        zfull[:, grp * per:(grp + 1) * per, :n, p * w:(p + 1) * w] = flat[r]
    res = finish_fn(zfull)
    return res[0] if squeeze else res


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
