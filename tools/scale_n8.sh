#!/bin/bash
# Developer tool (8-GPU box): configs[1] at 1 and 8 GPUs on the same box (weak scaling)
python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/r2f_scale_n1.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 100)) \
  bench.py --gpus 8 --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/r2f_scale_n8.log
python - <<PY
import json
a = json.loads(open("gpurun_out/r2f_scale_n1.log").read()); b = json.loads(open("gpurun_out/r2f_scale_n8.log").read())
print("N=1 ms/step", round(a["ms_per_step"], 3), "e2e", round(a["e2e"]["ms_per_step"], 3), "| N=8 ms/step", round(b["ms_per_step"], 3), "e2e", round(b["e2e"]["ms_per_step"], 3),
      "| efficiency", round(b["value"] / (8 * a["value"]), 4), "Mvox/s", round(b["value"] / 1e6, 1))
PY
