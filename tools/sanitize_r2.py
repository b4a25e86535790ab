"""Round-2 kernels for compute-sanitizer (memcheck and racecheck): fused iadrt sweeps (warp rings guarded by
__syncwarp), fused normal operator (R-layout hand-off), stitch / unstitch / truncate gathers,
truncate_mean over gathered shares, the angle-block part transforms and the co-scheduled passes.
  compute-sanitizer --tool memcheck  python tools/sanitize_r2.py
  compute-sanitizer --tool racecheck python tools/sanitize_r2.py race"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _adrt_cdefs as cd  # noqa: E402
from adrt_b200 import _lib  # noqa: E402
from adrt_b200._shard import M_LAST  # noqa: E402

race = len(sys.argv) > 1 and sys.argv[1] == "race"
lib = _lib.load()
for dt in (torch.float32, torch.float64):
    code = _lib.F32 if dt == torch.float32 else _lib.F64
    for n, B in ((4, 2), (64, 2), (256, 1), (1024, 1)) if not race else ((64, 1), (512, 1)):
        x = torch.rand((B, n, n), device="cuda", dtype=dt)
        y = adrt.adrt(x)
        for sp in (None, "5,1" if n == 64 else None):
            os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
            if sp:
                os.environ["ADRT_B200_IADRT_SPLIT"] = sp
            w = adrt.iadrt(y)
        os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
        no = cd.normal_operator(x)
        st = adrt.utils.stitch_adrt(y, remove_repeated=True)
        un = adrt.utils.unstitch_adrt(st)
        tr = adrt.utils.truncate(y)
        if n >= 64:
            # shares layout: 2 groups x 2 ranks
            shares = torch.stack([y[:, (r // 2) * 2:(r // 2) * 2 + 2, :n, (r % 2) * (n // 2):(r % 2 + 1) * (n // 2)].contiguous()
                                  for r in range(4)])
            got = torch.empty((B, n, n), device="cuda", dtype=dt)
            _lib.check(lib.adrt_b200_truncate_mean_shares(shares.data_ptr(), got.data_ptr(), B, n, 2, 2, 1.0, code,
                                                          torch.cuda.current_stream().cuda_stream), "shares")
        if n >= 16:
            parts, planes, D = 2, 4 * B, 2 * n - 1
            pf = int(lib.adrt_b200_part_exchange_pitch(n, code, M_LAST, 1))
            nb = int(lib.adrt_b200_part_workspace_bytes(planes, n, code, M_LAST))
            ws = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
            xf = torch.zeros((planes, 1 << M_LAST, n >> M_LAST, pf), dtype=dt, device="cuda")
            sino = torch.zeros((B, 4, D, n), dtype=dt, device="cuda")
            s = torch.cuda.current_stream().cuda_stream
            for p in range(parts):
                _lib.check(lib.adrt_b200_adrt_part(x.data_ptr(), xf.data_ptr(), None, B, n, code, 0, 4, p, parts, M_LAST, 0, ws.data_ptr(), nb, s), "part0")
            for p in range(parts):
                _lib.check(lib.adrt_b200_adrt_part(None, xf.data_ptr(), sino.data_ptr(), B, n, code, 0, 4, p, parts, M_LAST, 1, ws.data_ptr(), nb, s), "part1")
        torch.cuda.synchronize()
        print("ok", dt, n, B, float(w.abs().max()), float(no.sum()), float(un.sum()), float(tr.sum()), flush=True)
# co-scheduled passes (persistent kernels, release / acquire counters)
os.environ["ADRT_B200_COSCHED"] = "1"
for n in (1024,) if race else (1024, 2048):
    x = torch.rand((2, n, n), device="cuda")
    y = adrt.adrt(x)
    z = adrt.bdrt(y)
    torch.cuda.synchronize()
    print("ok cosched", n, float(z.sum()), flush=True)
