"""Staged passes (stage_tile.h: tensor-map / bulk copies in, bulk copies out, persistent CTAs, two buffers
swapping roles) for compute-sanitizer: both directions (the transposed variant is opt-in), interior and
boundary tiles, several tiles per CTA, checked against the default kernels.
  compute-sanitizer --tool memcheck  python tools/sanitize_staged.py
  compute-sanitizer --tool racecheck python tools/sanitize_staged.py race"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt  # noqa: E402
from adrt_b200 import _adrt_cdefs as cd  # noqa: E402

race = len(sys.argv) > 1 and sys.argv[1] == "race"
for n, B, split in ((64, 2, "5,1"), (512, 1, None)) if race else ((64, 3, "5,1"), (256, 2, "5,3"), (1024, 2, None), (2048, 3, None)):
    x = torch.randn((B, n, n), device="cuda")
    s = torch.randn((B, 4, 2 * n - 1, n), device="cuda")
    got = {}
    for tag, stage in (("base", ""), ("staged", "f5p,b5p")):
        os.environ["ADRT_B200_STAGE_SET"] = stage
        for k in ("ADRT_B200_SPLIT", "ADRT_B200_SPLIT_BDRT"):
            os.environ.pop(k, None)
        if split:
            os.environ["ADRT_B200_SPLIT"] = split
            os.environ["ADRT_B200_SPLIT_BDRT"] = ",".join(reversed(split.split(",")))
        got[tag] = (adrt.adrt(x), adrt.bdrt(s), cd.bdrt_planes(s, rows=n)[..., :n, :], cd.normal_operator(x, 1.0))
    torch.cuda.synchronize()
    same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(got["staged"], got["base"]))
    print("ok" if same else "MISMATCH", n, B, split, flush=True)
    assert same
