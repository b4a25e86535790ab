"""Batch-sharded inverses (iadrt and iadrt_fmg shard by batch only, "replicas" per image).
A global batch of B sinograms is split B / world per rank; every rank runs adrt.iadrt and one multigrid pass
(adrt.core.iadrt_fmg_step) on its shard, no collective on the data path.  Prints the max-over-ranks time,
the aggregate throughput and whether rank 0's shard equals the same images computed in one piece.
usage: [torchrun --nproc-per-node N] python tools/inverse_sharded.py [B n]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200._shard import max_over_ranks, shard_bounds  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d = dist if world > 1 else None
lo, hi = shard_bounds(B, world, rank)
g = torch.Generator(device=dev).manual_seed(5)           # same global batch on every rank, each keeps its shard
full = torch.rand((B, n, n), device=dev, generator=g)
y = adrt.adrt(full[lo:hi].contiguous())
del full


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    if d is not None:
        d.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / reps, d, dev), out


t_inv, inv = timed(lambda: adrt.iadrt(y))
t_fmg, fmg = timed(lambda: adrt.core.iadrt_fmg_step(y))
# the shard's first image on its own must give the same bytes (batch items never interact)
same = bool(torch.equal(adrt.iadrt(y[:1]).view(torch.int32), inv[:1].view(torch.int32)) and
            torch.equal(adrt.core.iadrt_fmg_step(y[:1]).view(torch.int32), fmg[:1].view(torch.int32)))
if rank == 0:
    print(json.dumps({"B": B, "n": n, "world": world, "images_per_rank": hi - lo,
                      "iadrt_ms": round(t_inv, 3), "iadrt_Gpixel_s": round(B * n * n / t_inv / 1e6, 2),
                      "fmg_step_ms": round(t_fmg, 3), "fmg_step_Gpixel_s": round(B * n * n / t_fmg / 1e6, 2),
                      "shard_equals_single": same}), flush=True)
if world > 1:
    dist.destroy_process_group()
