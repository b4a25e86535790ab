"""adrt + bdrt at sizes that use every fused pass kind, for compute-sanitizer --tool racecheck."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402

for dt, sizes in ((torch.float32, (64, 256, 1024, 2048, 4096)), (torch.float64, (64, 1024, 2048))):
    for n in sizes:
        x = torch.rand((1, n, n), device="cuda", dtype=dt)
        y = adrt.adrt(x)
        z = adrt.bdrt(y)
        torch.cuda.synchronize()
        print("ok", dt, n, float(z.sum()), flush=True)
