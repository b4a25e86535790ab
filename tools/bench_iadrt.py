"""iadrt timings (device resident, CUDA events, median of 5): fused multi-stage passes for several
stage splits and the one-kernel-per-stage path.  usage: python tools/bench_iadrt.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _lib  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


lib = _lib.load()
QUICK = os.environ.get("QUICK") == "1"   # default split only, two shapes
DEFAULTS = os.environ.get("DEFAULTS") == "1"   # every shape, default split only
for (B, n, dt, splits) in ((16, 2048, torch.float32, [None]), (8, 2048, torch.float64, [None])) if QUICK else ((16, 2048, torch.float32, [None, "4,4,3", "5,5,1", "1,5,5", "2,4,5", "3,4,4", "2,5,4"]),
                           (8, 2048, torch.float64, [None, "4,4,3", "3,4,4", "2,4,5"]),
                           (4, 4096, torch.float32, [None, "5,5,2", "4,4,4", "2,5,5"]),
                           (64, 1024, torch.float32, [None, "5,5", "3,4,3", "2,4,4"]),
                           (1, 8192, torch.float32, [None])):
    if DEFAULTS:
        splits = [None]
    y = torch.rand((B, 4, 2 * n - 1, n), device="cuda", dtype=dt)
    out = torch.empty_like(y)
    S = y.numel() * y.element_size()
    if not QUICK and not DEFAULTS:
        lib.adrt_b200_set_mode(1)
        t = timeit(lambda: adrt.iadrt(y, out=out))
        lib.adrt_b200_set_mode(0)
        print(json.dumps({"B": B, "n": n, "dtype": str(dt), "path": "per-stage", "iadrt_ms": round(t, 3)}), flush=True)
    for sp in splits:
        os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
        if sp:
            os.environ["ADRT_B200_IADRT_SPLIT"] = sp
        t = timeit(lambda: adrt.iadrt(y, out=out))
        print(json.dumps({"B": B, "n": n, "dtype": str(dt), "path": "fused", "split": sp or "default", "iadrt_ms": round(t, 3),
                          "compulsory_GBs": round(2 * S / (t * 1e-3) / 1e9, 1)}), flush=True)
    os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
    del y, out
    torch.cuda.empty_cache()
