#!/bin/bash
# ncu --set full capture of the pass kernels of ONE adrt + bdrt pair (B = 8, n = 2048) for fp32 and fp64
# (warm-up pair skipped), summarised on the box (gpurun_out/ carries at most 64 MiB back: the fp64 report
# is reduced to its text summaries, the fp32 report travels).
# usage (on the GPU box): bash tools/capture_head.sh <outdir under gpurun_out> <commit>
out=gpurun_out/$1; commit=$2; mkdir -p $out
K='regex:pass_kernel|stream_kernel|staged_kernel'
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 7 --launch-count 7 \
  -o $out/prof_f32 -f python tools/prof_once.py 8 2048 f32 > $out/ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 4 --launch-count 4 \
  -o $out/prof_f64 -f python tools/prof_once.py 8 2048 f64 > $out/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
for t in f32 f64; do
  python tools/ncu_summary.py $out/prof_$t.ncu-rep > $out/summary_$t.txt 2>&1
  python tools/ncu_traffic.py $out/prof_$t.ncu-rep 8 2048 $t $commit > $out/traffic_$t.json 2>$out/traffic_$t.err
  python tools/ncu_phases.py $out/prof_$t.ncu-rep > $out/phases_$t.txt 2>&1
done
ncu -i $out/prof_f64.ncu-rep --page source --csv > $out/source_f64.csv 2>/dev/null; gzip -9 $out/source_f64.csv
rm -f $out/prof_f64.ncu-rep
ls -la $out
