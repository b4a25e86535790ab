"""Tuning sweep on one GPU: time adrt and bdrt (device resident, CUDA events)
for a list of pass splits / wave sizes.  Usage: python tools/sweep.py [B n dtype]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dt = torch.float32 if (len(sys.argv) <= 3 or sys.argv[3] == "f32") else torch.float64
configs = json.loads(sys.argv[4]) if len(sys.argv) > 4 else [{}]
isz = 4 if dt == torch.float32 else 8
peak = 6466.1

x = torch.rand((B, n, n), device="cuda", dtype=dt)
y = torch.empty((B, 4, 2 * n - 1, n), device="cuda", dtype=dt)
z = torch.empty_like(y)
fb = (n * n + 4 * (2 * n - 1) * n) * isz * B
bb = 2 * 4 * (2 * n - 1) * n * isz * B


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


ref_y = ref_z = None
for cfg in configs:
    for k in [k for k in os.environ if k.startswith("ADRT_B200_") and k not in ("ADRT_B200_LIB", "ADRT_B200_DEVICE")]:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        if k == "mode":
            continue
        os.environ[k] = str(v)
    _lib.load().adrt_b200_set_mode(int(cfg.get("mode", 0)))
    ta = timeit(lambda: adrt.adrt(x, out=y))
    tb = timeit(lambda: adrt.bdrt(y, out=z))
    gpx = B * n * n / ((ta + tb) * 1e-3) / 1e9
    # every configuration must produce the bytes of the first one
    if ref_y is None:
        ref_y, ref_z = y.clone(), z.clone()
        same = True
    else:
        same = bool(torch.equal(y.view(torch.uint8), ref_y.view(torch.uint8)) and
                    torch.equal(z.view(torch.uint8), ref_z.view(torch.uint8)))
    print(json.dumps({"B": B, "n": n, "dtype": str(dt), "cfg": cfg, "same_bytes_as_first": same, "adrt_ms": round(ta, 3), "bdrt_ms": round(tb, 3),
                      "Gpx/s": round(gpx, 2), "adrt_frac": round(fb / (ta * 1e-3) / 1e9 / peak, 3),
                      "bdrt_frac": round(bb / (tb * 1e-3) / 1e9 / peak, 3),
                      "frac": round((fb + bb) / ((ta + tb) * 1e-3) / 1e9 / peak, 3)}), flush=True)
