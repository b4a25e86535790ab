"""Compact per-kernel register/spill table from a ptxas -v log."""
import re, sys
log = open(sys.argv[1]).read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", re.S)
rows = []
for m in pat.finditer(log):
    name = m.group(1)
    k = re.search(r"pass_kernelI([fd])Li(\d)ELi(\d)ELi(\d)ELb([01])", name)
    key = f"{k.group(1)} M{k.group(2)} L{k.group(3)} S{k.group(4)} {'fwd' if k.group(5)=='1' else 'bwd'}" if k else name[:40]
    rows.append((key, int(m.group(4)), int(m.group(3))))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for key, regs, spill in sorted(rows):
    if flt in key or spill:
        print(f"{key:28s} regs={regs:3d} spill={spill}")
