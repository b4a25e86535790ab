#!/bin/bash
# Round-2 multi-GPU measurements on ONE 8-GPU box (results under gpurun_out/):
#   * BASELINE config 3: 64 x 2048^2 fp64, batch sharded 64/G per GPU, G = 1, 2, 4, 8 (device resident)
#   * the headline fp32 workload at G = 8 (own check of the strong-scaled line the driver records)
#   * BASELINE config 5: one 8192^2 fp32 image, normal operator + CG iterations on 1 / 2 / 4 / 8 GPUs
#     (quadrants up to 4 ranks, 4 quadrants x 2 angle halves on 8), bit-compared with the 1-GPU result
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for G in 1 2 4 8; do
  if [ $G -eq 1 ]; then
    timeout 300 python bench.py --dtype f64 --steps 5 --no-e2e --no-cpu > $OUT/r02_bench_f64_n1.json 2> $OUT/r02_bench_f64_n1.err
  else
    timeout 300 $TR --nproc-per-node $G --master-port $((29600+G)) bench.py --gpus $G --dtype f64 --steps 5 --no-e2e --no-cpu > $OUT/r02_bench_f64_n$G.json 2> $OUT/r02_bench_f64_n$G.err
  fi
  tail -c 400 $OUT/r02_bench_f64_n$G.json; echo
done
timeout 300 $TR --nproc-per-node 8 --master-port 29650 bench.py --gpus 8 --steps 10 --no-cpu > $OUT/r02_bench_f32_n8.json 2> $OUT/r02_bench_f32_n8.err
head -c 600 $OUT/r02_bench_f32_n8.json; echo
: > $OUT/r02_cg_sharded.jsonl
timeout 200 python tools/cg_sharded.py 8192 4 2>&1 | grep "^{" | tee -a $OUT/r02_cg_sharded.jsonl
for G in 2 4 8; do
  timeout 300 $TR --nproc-per-node $G --master-port $((29700+G)) tools/cg_sharded.py 8192 4 2>&1 | grep -E "^\{|Error" | tee -a $OUT/r02_cg_sharded.jsonl
done
ADRT_B200_SHARD_PARTS=2 timeout 300 $TR --nproc-per-node 4 --master-port 29720 tools/cg_sharded.py 8192 4 2>&1 | grep -E "^\{|Error" | tee -a $OUT/r02_cg_sharded.jsonl
# inverses, batch sharded (16 x 2048^2 fp32 split over the ranks)
: > $OUT/r02_inverse_sharded.jsonl
timeout 200 python tools/inverse_sharded.py 16 2048 2>&1 | grep "^{" | tee -a $OUT/r02_inverse_sharded.jsonl
for G in 2 4 8; do
  timeout 300 $TR --nproc-per-node $G --master-port $((29800+G)) tools/inverse_sharded.py 16 2048 2>&1 | grep -E "^\{|Error" | tee -a $OUT/r02_inverse_sharded.jsonl
done
nvidia-smi topo -m > $OUT/r02_topo.txt 2>&1
