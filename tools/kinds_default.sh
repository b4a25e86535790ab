#!/bin/bash
# per-kernel durations of the default plan (ncu timing pass, B = 8, n = 2048): bash tools/kinds_default.sh [f32|f64]
DT=${1:-f32}
ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_once.py 8 2048 $DT 2>/dev/null \
    | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
rows=rows[1:]
for r in rows[len(rows)//2:]:
    k=r[ki].replace('adrt_b200::','').replace('(anonymous namespace)::','').replace('<unnamed>::','')
    print('  %8.1f us  %s' % (float(r[vi].replace(',',''))/1000.0, k[:110]))
"
