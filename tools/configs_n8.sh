#!/bin/bash
# Developer tool (8-GPU box): BASELINE configs[2..4] at 8 GPUs, one JSON line each into gpurun_out/
run() { # name, args...
  local name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29530 + RANDOM % 100)) \
    bench.py --gpus 8 "$@" 2>&1 | grep "^{" > gpurun_out/r2_${name}_n8.log
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_${name}_n8.log").read())
print("${name}", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "value", d["value"],
      "buildings/s", d.get("buildings_per_sec"))
PY
}
run scale_fp32_b1 --steps 30 --warmup 3 --no-cpu-baseline
run cfg3_bf16_b2 --steps 20 --warmup 3 --batch 2 --precision bf16 --no-cpu-baseline
run cfg4_fp32_b4 --steps 10 --warmup 3 --batch 4 --no-cpu-baseline
run cfg5_infer_2M --steps 10 --warmup 3 --mode infer --no-cpu-baseline
