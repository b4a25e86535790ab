"""Device-resident timings of the remaining section-8 operators (CUDA events, median of 5)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _adrt_cdefs as cd  # noqa: E402


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


def gbs(nbytes, ms):
    return round(nbytes / (ms * 1e-3) / 1e9, 1)


out = []
for (B, n, dt) in ((16, 2048, torch.float32), (4, 4096, torch.float32), (8, 2048, torch.float64)):
    isz = 4 if dt == torch.float32 else 8
    x = torch.rand((B, n, n), device="cuda", dtype=dt)
    y = adrt.adrt(x)
    S = y.numel() * isz
    r = {"B": B, "n": n, "dtype": str(dt)}
    t = timeit(lambda: adrt.iadrt(y)); r["iadrt_ms"] = round(t, 3); r["iadrt_GBs(2*K*S)"] = gbs(2 * S * (n.bit_length() - 1), t)
    t = timeit(lambda: adrt.core.adrt_step(y, 3)); r["adrt_step_ms"] = round(t, 3); r["adrt_step_GBs"] = gbs(2 * S, t)
    t = timeit(lambda: adrt.core.bdrt_step(y, 3)); r["bdrt_step_ms"] = round(t, 3); r["bdrt_step_GBs"] = gbs(2 * S, t)
    t = timeit(lambda: adrt.utils.interp_to_cart(y)); r["interp_ms"] = round(t, 3)
    t = timeit(lambda: cd.press_fmg_restriction(y)); r["restriction_ms"] = round(t, 3)
    t = timeit(lambda: cd.press_fmg_highpass(x)); r["highpass_ms"] = round(t, 3); r["highpass_GBs"] = gbs(2 * x.numel() * isz, t)
    t = timeit(lambda: cd.press_fmg_prolongation(x)); r["prolongation_ms"] = round(t, 3)
    t = timeit(lambda: adrt.core.iadrt_fmg_step(y), reps=3); r["fmg_step_ms"] = round(t, 3)
    t = timeit(lambda: cd.truncate_mean(y, 3.0)); r["truncate_mean_ms"] = round(t, 3)
    t = timeit(lambda: adrt.bdrt(y)); r["bdrt_ms"] = round(t, 3)
    t = timeit(lambda: cd.bdrt_planes(y, rows=n)); r["bdrt_rows_n_ms"] = round(t, 3)
    t = timeit(lambda: cd.truncate_mean(adrt.bdrt(adrt.adrt(x)), 1.0)); r["normal_op_unfused_ms"] = round(t, 3)
    t = timeit(lambda: cd.bdrt_truncate_mean(adrt.adrt(x), 1.0)); r["normal_op_r1_rows_ms"] = round(t, 3)
    from adrt_b200 import recipes; t = timeit(lambda: recipes.normal_operator(x)); r["normal_op_ms"] = round(t, 3)
    out.append(r)
    print(json.dumps(r), flush=True)
    del x, y
    torch.cuda.empty_cache()

# BASELINE config 4: iadrt_fmg workload, 16 images of 4096^2 fp32 -- one FMG pass, device resident
x = torch.rand((16, 4096, 4096), device="cuda")
y = adrt.adrt(x)
del x
t = timeit(lambda: adrt.core.iadrt_fmg_step(y), reps=3)
print(json.dumps({"B": 16, "n": 4096, "dtype": "torch.float32", "fmg_step_ms": round(t, 3),
                  "Gpx/s": round(16 * 4096 * 4096 / (t * 1e-3) / 1e9, 2)}), flush=True)
