"""Per-phase (barrier-delimited) dynamic instruction breakdown of one kernel launch in an .ncu-rep.
usage: python tools/ncu_segments.py rep launch_index"""
import csv, re, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:pass_kernel",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 10]; data = data[:len(data) // 2]
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); iss = hdr.index("# Samples")
def I(r, k):
    try: return int(r[k])
    except ValueError: return 0
tot = sum(I(r, ia) for r in data); ts = sum(I(r, iss) for r in data)
print(rows[0][1][:100]); print("total warp-instr", tot, "static", len(data), "samples", ts)
seg = []; cur = [0, 0, 0, {}]
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc]); o = m.group(2) if m else '?'
    cur[0] += I(r, ia); cur[1] += I(r, iss); cur[2] += 1
    k = o.split('.')[0]
    if k in ('LDG', 'STG', 'LDS', 'STS', 'FADD', 'DADD'): cur[3][k] = cur[3].get(k, 0) + I(r, ia)
    if o.startswith('BAR'):
        seg.append(cur); cur = [0, 0, 0, {}]
seg.append(cur)
for i, c in enumerate(seg):
    print("seg%d static=%d exec=%.1f%% stall-samples=%.1f%% %s" % (i, c[2], 100 * c[0] / tot, 100 * c[1] / ts,
          {k: "%.1f%%" % (100 * v / tot) for k, v in c[3].items()}))
