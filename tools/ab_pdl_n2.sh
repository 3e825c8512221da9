#!/bin/bash
# Developer tool (2-GPU box): programmatic dependent launch off / on at 1 and 2 GPUs (gradient all-reduce in the step)
for m in 0 1; do
  for n in 1 2; do
    if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 100)) bench.py --gpus $n"; fi
    SCN_B200_PDL=$m $cmd --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PDL=$m N=$n ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3))"
  done
done
