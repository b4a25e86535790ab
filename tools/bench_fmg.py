"""iadrt_fmg_step timings (device resident, CUDA events, median of 5) with the residual subtracted by the
loader of bdrt's first pass (default) and by a kernel of its own (ADRT_B200_FMG_SUB_SEPARATE=1)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402


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


for (B, n, dt) in ((16, 2048, torch.float32), (4, 4096, torch.float32), (16, 4096, torch.float32), (8, 2048, torch.float64),
                   (64, 1024, torch.float32)):
    y = torch.rand((B, 4, 2 * n - 1, n), device="cuda", dtype=dt)
    r = {"B": B, "n": n, "dtype": str(dt)}
    for key, env in (("fmg_step_ms", None), ("fmg_step_separate_sub_ms", "1")):
        os.environ.pop("ADRT_B200_FMG_SUB_SEPARATE", None)
        if env:
            os.environ["ADRT_B200_FMG_SUB_SEPARATE"] = env
        r[key] = round(timeit(lambda: adrt.core.iadrt_fmg_step(y)), 3)
    os.environ.pop("ADRT_B200_FMG_SUB_SEPARATE", None)
    r["Gpx/s"] = round(B * n * n / (r["fmg_step_ms"] * 1e-3) / 1e9, 2)
    print(json.dumps(r), flush=True)
    del y
    torch.cuda.empty_cache()
