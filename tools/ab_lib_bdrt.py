"""bdrt (and adrt) of B x n^2 fp32 with the library ADRT_B200_LIB names: median of 7 (CUDA events) and a checksum of
the result for a bytes-equal comparison across processes.  usage: ADRT_B200_LIB=... python tools/ab_lib_bdrt.py [B n]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.rand((B, 4, 2 * n - 1, n), device="cuda", dtype=torch.float32, generator=g)
z = torch.empty_like(y)
for _ in range(3):
    adrt.bdrt(y, out=z)
torch.cuda.synchronize()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); adrt.bdrt(y, out=z); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
chk = int(z.view(torch.int32).to(torch.int64).sum().item())
print(json.dumps({"lib": os.path.basename(os.environ.get("ADRT_B200_LIB", "default")), "B": B, "n": n,
                  "bdrt_ms": round(ts[3], 3), "min_ms": round(ts[0], 3), "checksum": chk}), flush=True)
