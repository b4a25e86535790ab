"""Instruction histogram per kernel of a cuobjdump -sass listing (tools/sass_hist.py file.sass [filter])."""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if flt not in name:
        continue
    ops = collections.Counter()
    for line in f.split('\n'):
        mm = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if mm:
            ops[mm.group(1)] += 1
    print(name[:150])
    print("  total", sum(ops.values()), dict(ops.most_common(24)))
