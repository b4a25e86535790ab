"""BASELINE config 5 style run: CG reconstruction of ONE large image with the normal
operator sharded over the ranks (torchrun, NCCL) -- by quadrant up to 4 ranks, by quadrant x
angle block beyond (ADRT_B200_SHARD_PARTS forces the angle-block path on fewer ranks).  Prints ms
per operator application and per CG iteration (max over ranks) and checks the sharded results
against the 1-GPU ones bit for bit.
usage: torchrun --nproc-per-node N tools/cg_sharded.py [n] [iters]"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import recipes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(0)  # same image on every rank
ys, xs = torch.meshgrid(torch.linspace(-1, 1, n, device=dev), torch.linspace(-1, 1, n, device=dev), indexing="ij")
img = torch.exp(-8 * (xs ** 2 + (ys - 0.3) ** 2)) + 0.5 * torch.exp(-30 * ((xs + 0.4) ** 2 + ys ** 2))
b = adrt.adrt(img)
d = dist if world > 1 else None
# warm-up + timing of the operator alone
for _ in range(2):
    recipes.normal_operator(img, dist=d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    out = recipes.normal_operator(img, dist=d)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
single = recipes.normal_operator(img)  # every rank can also do the whole thing alone
same = bool(torch.equal(out.view(torch.int32), single.view(torch.int32)))
# a fixed number of CG iterations on the normal equations (docs/examples.cginverse.md:40-67), sharded vs alone
from adrt_b200._shard import image_layout  # noqa: E402


def cg_fixed(its, dd):
    try:
        recipes.iadrt_cg(b, maxiter=its, rtol=1e-30, dist=dd)
    except ValueError:      # "convergence failed": expected, the iteration count is the point
        pass
    torch.cuda.synchronize()


cg_fixed(1, d)
torch.cuda.synchronize()
t0 = time.perf_counter()
cg_fixed(iters, d)
cg_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / iters], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(cg_ms, op=dist.ReduceOp.MAX)
per, parts = image_layout(world)
if rank == 0:
    import json
    print(json.dumps({"n": n, "world": world, "quadrants_per_group": per, "ranks_per_group": parts,
                      "normal_operator_ms": round(float(ms.item()), 3), "cg_iteration_ms": round(float(cg_ms.item()), 3),
                      "bit_identical_to_1gpu": same}), flush=True)
if world > 1:
    dist.destroy_process_group()
