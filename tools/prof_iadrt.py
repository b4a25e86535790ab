"""One warm-up + one measured iadrt, for ncu captures.  Usage: python tools/prof_iadrt.py [B n dtype]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dt = torch.float32 if (len(sys.argv) <= 3 or sys.argv[3] == "f32") else torch.float64
y = torch.rand((B, 4, 2 * n - 1, n), device="cuda", dtype=dt)
out = torch.empty_like(y)
for _ in range(2):
    adrt.iadrt(y, out=out)
    torch.cuda.synchronize()
