"""Sharded normal operator of ONE n x n image (BASELINE config 5) on the ranks of this torchrun job: the slab
exchange (ADRT_B200_SHARD_GATHER=0, default from 4 ranks on) and the all-gather form (=1) timed back to back in the same
processes (CUDA events, max over ranks), each checked bit for bit against the 1-GPU operator; then CG
iterations with the default.  One JSON line per mode from rank 0.
usage: torchrun --nproc-per-node N tools/normal_op_sharded_ab.py [n] [iters]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import recipes  # noqa: E402
from adrt_b200._shard import image_layout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ys, xs = torch.meshgrid(torch.linspace(-1, 1, n, device=dev), torch.linspace(-1, 1, n, device=dev), indexing="ij")
img = torch.exp(-8 * (xs ** 2 + (ys - 0.3) ** 2)) + 0.5 * torch.exp(-30 * ((xs + 0.4) ** 2 + ys ** 2))
del ys, xs
d = dist if world > 1 else None
single = recipes.normal_operator(img)
per, parts = image_layout(world)
for mode in (("slab", "0"), ("gather", "1")) if world > 1 else (("single", "0"),):
    os.environ["ADRT_B200_SHARD_GATHER"] = mode[1]
    for _ in range(2):
        out = recipes.normal_operator(img, dist=d)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = recipes.normal_operator(img, dist=d)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(out.view(torch.int32), single.view(torch.int32)))
    if rank == 0:
        print(json.dumps({"n": n, "world": world, "quadrants_per_group": per, "ranks_per_group": parts, "mode": mode[0],
                          "normal_operator_ms": round(float(ms.item()), 3), "bit_identical_to_1gpu": same}), flush=True)
os.environ.pop("ADRT_B200_SHARD_GATHER", None)      # the default form for this world size
b = adrt.adrt(img)


def cg_fixed(its):
    try:
        recipes.iadrt_cg(b, maxiter=its, rtol=1e-30, dist=d)
    except ValueError:      # "convergence failed": expected, the iteration count is the point
        pass
    torch.cuda.synchronize()


cg_fixed(1)
t0 = time.perf_counter()
cg_fixed(iters)
cg_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / iters], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(cg_ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"n": n, "world": world, "mode": "cg, default exchange", "cg_iteration_ms": round(float(cg_ms.item()), 3)}), flush=True)
if world > 1:
    dist.destroy_process_group()
