"""iadrt timings (device resident, CUDA events, median of 5) for a few shapes."""
import json, os, sys
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

for (B, n, dt) in ((16, 2048, torch.float32), (4, 4096, torch.float32), (1, 4096, torch.float32), (8, 2048, torch.float64), (64, 512, torch.float32)):
    y = torch.rand((B, 4, 2 * n - 1, n), device="cuda", dtype=dt)
    out = torch.empty_like(y)
    t = timeit(lambda: adrt.iadrt(y, out=out))
    S = y.numel() * y.element_size()
    print(json.dumps({"B": B, "n": n, "dtype": str(dt), "batch": os.environ.get("ADRT_B200_IADRT_BATCH", "8"), "iadrt_ms": round(t, 3),
                      "GB/s(2*K*S)": round(2 * S * (n.bit_length() - 1) / t / 1e6, 1)}), flush=True)
    del y, out
