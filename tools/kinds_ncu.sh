#!/bin/bash
# per-kernel durations (ncu timing pass, B = 8, n = 2048 fp32) for alternative pass splits / kernel kinds
# usage (on the GPU box): bash tools/kinds_ncu.sh > gpurun_out/kinds.txt
run() {
  echo "== $*"
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_once.py 8 2048 f32 2>/dev/null \
    | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
rows=rows[1:]
half=rows[len(rows)//2:]
for r in half:
    k=r[ki].replace('adrt_b200::','').replace('(anonymous namespace)::','').replace('<unnamed>::','')
    print('  %8.1f us  %s' % (float(r[vi].replace(',',''))/1000.0, k[:110]))
"
}
run A=1
run ADRT_B200_STREAM_SET=all
run ADRT_B200_SPLIT=6,5
run ADRT_B200_SPLIT=6,5 ADRT_B200_STREAM_SET=all
run ADRT_B200_SPLIT_BDRT=5,6
run ADRT_B200_SPLIT_BDRT=5,6 ADRT_B200_STREAM_SET=all
run ADRT_B200_STREAM_SET=none
