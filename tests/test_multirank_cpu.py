"""CPU, world_size 2 over gloo: the batch-sharding host logic used for multi-GPU
runs (shard bounds, max-over-ranks timing, result gathering).  The per-rank
transform is the CPU oracle here -- the test is about the sharding, which is
the same code path bench.py drives with NCCL on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from adrt_b200._shard import gather_batch, max_over_ranks, shard_bounds
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, n = 5, 16  # uneven split on purpose
        x = np.random.default_rng(0).standard_normal((B, n, n)).astype(np.float32)
        lo, hi = shard_bounds(B, world, rank)
        local = torch.from_numpy(O.adrt(x[lo:hi]))
        full = gather_batch(local, dist)
        ok = full.numpy().tobytes() == O.adrt(x).tobytes()
        # the inverses shard the same way (batch replicas, no exchange): exact inverse and
        # one multigrid pass of this rank's sinograms, gathered, equal the unsharded results
        y = O.adrt(x)
        inv = gather_batch(torch.from_numpy(O.iadrt(y[lo:hi])), dist)
        fmg = gather_batch(torch.from_numpy(O.iadrt_fmg_step(y[lo:hi])), dist)
        ok = ok and inv.numpy().tobytes() == O.iadrt(y).tobytes() and fmg.numpy().tobytes() == O.iadrt_fmg_step(y).tobytes()
        t = max_over_ranks(10.0 + rank, dist)
        q.put((rank, lo, hi, ok, t))
    finally:
        dist.destroy_process_group()


def _worker_quadrants(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from adrt_b200._shard import quadrant_owner_range, sharded_normal_operator
    from oracle import oracle as O

    def oracle_fn(x, q_first, q_count, part, parts, base_rank, dist_):
        # CPU stand-in for the rank's share: full oracle transform, keep our quadrants / columns
        n = x.shape[-1]
        z = torch.from_numpy(O.bdrt(O.adrt(x.numpy())))
        w = n // parts
        return z[:, q_first:q_first + q_count, :n, part * w:(part + 1) * w].contiguous()

    def oracle_finish(zfull):
        t = O.truncate(zfull.numpy())
        return torch.from_numpy((((t[:, 0] + t[:, 1]) + t[:, 2]) + t[:, 3]) / 4)

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 16
        x = torch.from_numpy(np.random.default_rng(3).standard_normal((n, n)))
        got = sharded_normal_operator(x, dist, local_fn=oracle_fn, finish_fn=oracle_finish).numpy()
        # the same with the two ranks sharing every quadrant as angle halves (columns of the back-projection)
        got2 = sharded_normal_operator(x, dist, local_fn=oracle_fn, finish_fn=oracle_finish, parts=2).numpy()
        # default finish: slab exchange + ordered local sum + all-gather of the slabs (_slab_mean), both layouts,
        # single image and batch
        y = O.bdrt(O.adrt(x.numpy()))
        t = O.truncate(y)
        xb3 = torch.from_numpy(np.random.default_rng(4).standard_normal((3, n, n)))
        yb = O.bdrt(O.adrt(xb3.numpy()))
        tb = O.truncate(yb)
        wantb = (((tb[:, 0] + tb[:, 1]) + tb[:, 2]) + tb[:, 3]) / 4
        ok_slab = True
        os.environ["ADRT_B200_SHARD_GATHER"] = "0"      # the slab form is the default from 4 ranks on; force it for 2
        for pp in (1, 2):
            g1 = sharded_normal_operator(x, dist, local_fn=oracle_fn, parts=pp).numpy()
            g3 = sharded_normal_operator(xb3, dist, local_fn=oracle_fn, parts=pp).numpy()
            ok_slab = ok_slab and g1.tobytes() == ((((t[0] + t[1]) + t[2]) + t[3]) / 4).tobytes() and g3.tobytes() == wantb.tobytes()
        os.environ.pop("ADRT_B200_SHARD_GATHER", None)
        y = O.bdrt(O.adrt(x.numpy()))
        t = O.truncate(y)
        want = (((t[0] + t[1]) + t[2]) + t[3]) / 4
        q.put((rank, quadrant_owner_range(world, rank),
               got.tobytes() == want.tobytes() and got2.tobytes() == want.tobytes() and ok_slab))
    finally:
        dist.destroy_process_group()


def test_two_rank_quadrant_sharded_normal_operator():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_quadrants, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 2), (2, 2)]
    assert all(r[2] for r in res), "sharded normal operator differs from the single-process one"


def test_quadrant_owner_range():
    from adrt_b200._shard import quadrant_owner_range

    assert quadrant_owner_range(1, 0) == (0, 4)
    assert [quadrant_owner_range(2, r) for r in range(2)] == [(0, 2), (2, 2)]
    assert [quadrant_owner_range(4, r) for r in range(4)] == [(0, 1), (1, 1), (2, 1), (3, 1)]
    # 8 ranks: 4 quadrants x 2 angle halves (every rank works)
    assert [quadrant_owner_range(8, r) for r in range(8)] == [(0, 1), (0, 1), (1, 1), (1, 1), (2, 1), (2, 1), (3, 1), (3, 1)]
    from adrt_b200._shard import image_layout, part_of

    assert [part_of(8, r) for r in range(8)] == [(r % 2, 2) for r in range(8)]
    assert image_layout(16) == (1, 4) and image_layout(2) == (2, 1)
    # explicit group size: 2 ranks sharing every quadrant as two angle halves
    assert image_layout(2, 2) == (4, 2) and quadrant_owner_range(2, 1, 2) == (0, 4) and part_of(2, 1, 2) == (1, 2)
    assert image_layout(8, 4) == (2, 4) and quadrant_owner_range(8, 5, 4) == (2, 2)
    with pytest.raises(ValueError, match="supports world sizes"):
        image_layout(3)
    with pytest.raises(ValueError, match="supports world sizes"):
        quadrant_owner_range(6, 0)
    with pytest.raises(ValueError, match="cannot split"):
        image_layout(32, 2)


def _worker_exchange(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from adrt_b200._shard import exchange_rows

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        planes, nblk, e, pitch = 2, 4, 6, 5
        # every cell carries (plane, blk, a, x) encoded as a number; cells a rank does not own start as -1
        full = (torch.arange(planes * nblk * e * pitch, dtype=torch.float64).reshape(planes, nblk, e, pitch))
        bl = slice(rank * nblk // world, (rank + 1) * nblk // world)
        an = slice(rank * e // world, (rank + 1) * e // world)
        for forward in (True, False):
            buf = torch.full_like(full, -1.0)
            if forward:
                buf[:, bl] = full[:, bl]          # all angles of my blocks
            else:
                buf[:, :, an] = full[:, :, an]    # all blocks of my angles
            exchange_rows(buf, rank, world, 0, dist, forward)
            want = torch.full_like(full, -1.0)
            if forward:
                want[:, bl] = full[:, bl]
                want[:, :, an] = full[:, :, an]   # now also every block of my angles
            else:
                want[:, :, an] = full[:, :, an]
                want[:, bl] = full[:, bl]         # now also every angle of my blocks
            ok = ok and bool(torch.equal(buf, want))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_angle_block_exchange():
    """The row exchange of angle-block sharding (adrt_b200/_shard.exchange_rows), world size 2 over gloo:
    after it a rank holds every block of its angles (forward) / every angle of its blocks (transposed),
    and nothing it was not sent."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "exchange_rows left the wrong rows behind"


def test_shard_bounds():
    from adrt_b200._shard import shard_bounds

    for B in (1, 5, 64, 67):
        for world in (1, 2, 4, 8):
            cuts = [shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_two_rank_gloo_sharded_transform():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[3] for r in res), "gathered sharded result differs from the unsharded one"
    assert all(r[4] == 11.0 for r in res), "max over ranks"
