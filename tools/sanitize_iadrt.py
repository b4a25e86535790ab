"""Fused iadrt sweeps (double-mapped warp rings) for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_iadrt.py
  compute-sanitizer --tool racecheck python tools/sanitize_iadrt.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402

for dt in (torch.float32, torch.float64):
    for n, sp in ((8, None), (32, None), (64, "5,1"), (64, None), (256, None), (512, "5,4"), (1024, None)):
        os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
        if sp:
            os.environ["ADRT_B200_IADRT_SPLIT"] = sp
        y = torch.rand((1, 4, 2 * n - 1, n), device="cuda", dtype=dt)
        w = adrt.iadrt(y)
        torch.cuda.synchronize()
        print("ok", dt, n, sp, float(w.abs().max()), flush=True)
os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
