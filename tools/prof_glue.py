"""One call of each device glue operator (for ncu captures): python tools/prof_glue.py [B n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _adrt_cdefs as cd  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
x = torch.rand((B, n, n), device="cuda")
y = adrt.adrt(x)
for _ in range(2):
    cd.truncate_mean(y, 3.0)
    cd.press_fmg_highpass(x)
    cd.press_fmg_restriction(y)
    cd.press_fmg_prolongation(x)
    adrt.core.bdrt_step(y, 3)
    adrt.utils.interp_to_cart(y)
    torch.cuda.synchronize()
