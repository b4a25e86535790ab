"""A/B of an environment switch read per call (e.g. ADRT_B200_STREAM_SPLIT2): device-resident adrt / bdrt timings
(CUDA events, median of 7) and a bytes-equal check against the unset default.
usage: python tools/bench_env_ab.py VAR v1,v2,... [BxN ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402


def timeit(fn, reps=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


var, vals = sys.argv[1], sys.argv[2].split(",")
dt = torch.float64 if os.environ.get("BENCH_DTYPE") == "f64" else torch.float32
iv = torch.int64 if dt == torch.float64 else torch.int32
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[3:]] or [(64, 2048)]
for (B, n) in shapes:
    x = torch.rand((B, n, n), device="cuda", dtype=dt)
    os.environ.pop(var, None)
    y0 = adrt.adrt(x)
    z0 = adrt.bdrt(y0)
    for v in [None] + vals:
        if v is None:
            os.environ.pop(var, None)
        else:
            os.environ[var] = v
        y = adrt.adrt(x)
        z = adrt.bdrt(y0)
        ok = bool(torch.equal(y.view(iv), y0.view(iv)) and torch.equal(z.view(iv), z0.view(iv)))
        del y, z
        ta = timeit(lambda: adrt.adrt(x))
        tb = timeit(lambda: adrt.bdrt(y0))
        print(json.dumps({"B": B, "n": n, "dtype": str(dt), var: v, "adrt_ms": round(ta, 3), "bdrt_ms": round(tb, 3),
                          "step_ms": round(ta + tb, 3), "bytes_equal_default": ok}), flush=True)
    del x, y0, z0
    torch.cuda.empty_cache()
