#!/bin/bash
# Developer tool (GPU box): the BASELINE configurations on one GPU, one JSON line each into gpurun_out/
run() { # name, args...
  local name=$1; shift
  python bench.py "$@" 2>&1 | grep "^{" > gpurun_out/r2f_${name}_n1.log
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2f_${name}_n1.log").read())
print("${name}", "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "inline", round(d["value_inline"]["ms_per_step"], 3),
      "pruned", round(d["value_pruned"]["ms_per_step"], 3), "Mvox/s", round(d["value"] / 1e6, 1),
      {k: round(v["ms_per_step"], 2) for k, v in d["kernel_classes"].items()}, "frac", round(d["roofline"]["frac"], 3))
PY
}
run tf32_b1 --steps 20 --warmup 3 --precision tf32 --no-cpu-baseline
run bf16_b1 --steps 20 --warmup 3 --precision bf16 --no-cpu-baseline
run cfg3_bf16_b2 --steps 20 --warmup 3 --batch 2 --precision bf16 --no-cpu-baseline
run cfg4_fp32_b4 --steps 10 --warmup 3 --batch 4 --no-cpu-baseline
run cfg5_infer_2M --steps 10 --warmup 3 --mode infer --no-cpu-baseline
