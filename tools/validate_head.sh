#!/bin/bash
# fp64 bench first (the number this change is about), then the whole GPU suite, smoke and the default bench
# usage (on the GPU box): bash tools/validate_head.sh <outdir under gpurun_out>
out=gpurun_out/$1; mkdir -p $out
python bench.py --dtype f64 --steps 5 --warmup 3 --no-e2e --no-cpu > $out/bench_f64_n1.json 2> $out/bench_f64_n1.err; echo "bench f64 rc=$?"
python -c "import json;d=json.load(open('$out/bench_f64_n1.json'));print('f64',d['ms_per_step'],d['roofline']['frac'],d['roofline']['split'])"
python bench.py --dtype f64 --n 4096 --batch 8 --steps 5 --warmup 3 --no-e2e --no-cpu > $out/bench_f64_4096.json 2> $out/bench_f64_4096.err
python -c "import json;d=json.load(open('$out/bench_f64_4096.json'));print('f64 4096',d['ms_per_step'],d['roofline']['split'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pass_kernel --launch-skip 4 --launch-count 4 --csv \
  python tools/prof_once.py 8 2048 f64 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print('  %8.1f us  %s' % (float(r[vi].replace(',',''))/1000.0, r[ki][:100]))
" | tee $out/kernels_f64.txt
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$out/bench_n1.json'));print('f32',d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['roofline']['traffic_source'])"
