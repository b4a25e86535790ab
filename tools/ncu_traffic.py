"""profiles/rNN_traffic.json from an ncu --set full capture of one adrt + bdrt pair
(tools/prof_once.py B n dtype): python tools/ncu_traffic.py rep B n dtype [commit] > profiles/rNN_traffic.json
bench.py reads the newest such file for roofline.traffic and reports the commit it was taken at."""
import csv
import json
import subprocess
import sys

rep, B, n, dtype = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
commit = sys.argv[5] if len(sys.argv) > 5 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit]


kernels, total = [], 0.0
for r in data:
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    total += rd + wr
    kernels.append({
        "kernel": r[ix["Kernel Name"]].replace("void adrt_b200::<unnamed>::", "").replace("adrt_b200::", ""),
        "dram_read_bytes_per_image": rd / B, "dram_write_bytes_per_image": wr / B,
        "duration_us_at_B%d" % B: to_us(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]]),
    })
itemsize = 4 if dtype == "f32" else 8
print(json.dumps({
    "capture": f"{rep} (ncu --set full --clock-control none, tools/prof_once.py {B} {n} {dtype})",
    "batch_in_capture": B, "B": B, "commit": commit, "n": n, "dtype": dtype, "kernels": kernels,
    "dram_bytes_per_image_fwd_plus_bdrt": total / B,
    "algorithmic_bytes_per_image": (25 * n * n - 12 * n) * itemsize,
}, indent=1))
