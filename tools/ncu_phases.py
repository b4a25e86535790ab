"""Stall samples of a kernel bucketed by the code regions between its BAR.SYNC instructions
(the tile programs are barrier-separated phases, so this is a per-phase time profile).
python tools/ncu_phases.py prof.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
seen = set()
for b in blocks[1:]:
    lines = b.split("\n")
    name = lines[0]
    if flt and flt not in name:
        continue
    if b in seen:   # the source page repeats a kernel's table once per view
        continue
    seen.add(b)
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    print(name[:160])
    seg, segs = {"n": 0, "samples": 0, "inst": 0, "wf": 0, "wf_ideal": 0, "stalls": {}}, []
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        s = int(r[ix["# Samples"]] or 0)
        seg["n"] += 1
        seg["samples"] += s
        seg["inst"] += int(r[ix["Instructions Executed"]] or 0)
        seg["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
        seg["wf_ideal"] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        for h in stall_cols:
            v = int(r[ix[h]] or 0)
            if v:
                seg["stalls"][h] = seg["stalls"].get(h, 0) + v
        if "BAR.SYNC" in r[ix["Source"]] or "EXIT" in r[ix["Source"]]:
            segs.append(seg)
            seg = {"n": 0, "samples": 0, "inst": 0, "wf": 0, "wf_ideal": 0, "stalls": {}}
    segs.append(seg)
    tot = sum(s["samples"] for s in segs) or 1
    for i, s in enumerate(segs):
        if s["samples"] == 0 and s["inst"] == 0:
            continue
        top = sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:4]
        print(f"  region {i}: sass {s['n']:5d} samples {100.0 * s['samples'] / tot:5.1f}% warp-inst {s['inst']:.3g} "
              f"smem wavefronts {s['wf']:.3g} (ideal {s['wf_ideal']:.3g})  " + " ".join(f"{k[6:]}={v}" for k, v in top))
