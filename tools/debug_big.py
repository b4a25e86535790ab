import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adrt_b200 as adrt
from adrt_b200 import _lib
lib = _lib.load()
def diff(a, b, what):
    ai, bi = a.view(torch.int32), b.view(torch.int32)
    ne = (ai != bi)
    cnt = int(ne.sum())
    msg = f"{what}: {cnt} of {a.numel()} differ"
    if cnt:
        idx = ne.nonzero()[0].tolist()
        idx2 = ne.nonzero()[-1].tolist()
        msg += f" first {idx} got {a[tuple(idx)].item()} want {b[tuple(idx)].item()} last {idx2}"
        # which quadrants / columns / rows are affected
        q = ne.any(dim=(0, 2, 3)).tolist(); msg += f" quadrants {q}"
        cols = ne.any(dim=(0, 1, 2)).nonzero().flatten(); msg += f" cols[{int(cols.min())}..{int(cols.max())}] n={cols.numel()}"
        rows = ne.any(dim=(0, 1, 3)).nonzero().flatten(); msg += f" rows[{int(rows.min())}..{int(rows.max())}] n={rows.numel()}"
    print(msg, flush=True)
for n, split in ((4096, "4,4,4"), (8192, "5,4,4"), (8192, "4,4,5"), (8192, "6,6,1")):
    os.environ["ADRT_B200_SPLIT"] = split; os.environ["ADRT_B200_SPLIT_BDRT"] = split
    x = torch.randn((1, n, n), device="cuda")
    lib.adrt_b200_set_mode(1); yr = adrt.adrt(x); zr = adrt.bdrt(yr)
    lib.adrt_b200_set_mode(0); y = adrt.adrt(x); z = adrt.bdrt(yr)
    diff(y, yr, f"adrt n={n} split={split}")
    diff(z, zr, f"bdrt n={n} split={split}")
    del x, y, z, yr, zr; torch.cuda.empty_cache()
