#!/bin/bash
# Round-2 multi-GPU measurements, part 2 (one 8-GPU box): the sharded 8192^2 normal operator / CG with the slab
# exchange (default) and with the older gather-everything form, and the batch-sharded inverses.
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > $OUT/r02_cg_sharded_slab.jsonl
timeout 200 python tools/cg_sharded.py 8192 4 2>&1 | grep "^{" | tee -a $OUT/r02_cg_sharded_slab.jsonl
for G in 2 4 8; do
  timeout 300 $TR --nproc-per-node $G --master-port $((29700+G)) tools/cg_sharded.py 8192 4 2>&1 | grep -E "^\{|Error" | tee -a $OUT/r02_cg_sharded_slab.jsonl
done
: > $OUT/r02_cg_sharded_gather.jsonl
for G in 4 8; do
  ADRT_B200_SHARD_GATHER=1 timeout 300 $TR --nproc-per-node $G --master-port $((29710+G)) tools/cg_sharded.py 8192 4 2>&1 | grep -E "^\{|Error" | tee -a $OUT/r02_cg_sharded_gather.jsonl
done
: > $OUT/r02_inverse_sharded.jsonl
timeout 200 python tools/inverse_sharded.py 16 2048 2>&1 | grep "^{" | tee -a $OUT/r02_inverse_sharded.jsonl
for G in 2 4 8; do
  timeout 300 $TR --nproc-per-node $G --master-port $((29800+G)) tools/inverse_sharded.py 16 2048 2>&1 | grep -E "^\{|Error" | tee -a $OUT/r02_inverse_sharded.jsonl
done
