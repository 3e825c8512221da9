#!/bin/bash
# Developer tool (8-GPU box): the gather-GEMM's static first work item on / off with the gradient all-reduce running
for m in 1 0 1 0; do
  SCN_B200_GEMM_STATIC_FIRST=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 100)) \
    bench.py --gpus 8 --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('static_first=$m N=8 ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3))"
done
