"""Small transforms of every kernel family for compute-sanitizer runs:
compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _adrt_cdefs as cd  # noqa: E402

rng = np.random.default_rng(0)
for dt in (torch.float32, torch.float64):
    for n, B in ((8, 2), (64, 2), (256, 1), (512, 1), (1024, 1), (2048, 1)):
        if dt == torch.float64 and n > 1024:
            continue
        x = torch.rand((B, n, n), device="cuda", dtype=dt)
        y = adrt.adrt(x)
        z = adrt.bdrt(y)
        w = adrt.iadrt(y)
        if n <= 512:
            for i in range(n.bit_length() - 1):
                adrt.core.adrt_step(y, i)
                adrt.core.bdrt_step(y, i)
        adrt.core.iadrt_fmg_step(y)
        cd.bdrt_planes(y, rows=n)
        adrt.utils.interp_to_cart(y)
        torch.cuda.synchronize()
        print("ok", dt, n, B, float(z.sum()), float(w.abs().max()), flush=True)
