#!/bin/bash
# ncu --set full capture of the four fp64 pass kernels of one adrt + bdrt pair (B = 8, n = 2048), text summaries only
# usage (on the GPU box): bash tools/capture_f64.sh <outdir under gpurun_out> <commit>
out=gpurun_out/$1; commit=$2; mkdir -p $out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:pass_kernel --launch-skip 4 --launch-count 4 \
  -o /tmp/prof_f64 -f python tools/prof_once.py 8 2048 f64 > $out/ncu_f64.log 2>&1; echo "ncu f64 rc=$?"
python tools/ncu_summary.py /tmp/prof_f64.ncu-rep > $out/summary_f64.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_f64.ncu-rep 8 2048 f64 $commit > $out/traffic_f64.json 2>/dev/null
python tools/ncu_phases.py /tmp/prof_f64.ncu-rep > $out/phases_f64.txt 2>&1
ls -la $out
