#!/bin/bash
# Developer tool (GPU box): the step with programmatic dependent launch off / on / on except the tensor-core GEMMs
for m in 0 1 2; do
  SCN_B200_PDL=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" > gpurun_out/pdl_$m.log
  python - <<PY
import json
d = json.loads(open("gpurun_out/pdl_$m.log").read())
print("PDL=$m ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "inline", round(d["value_inline"]["ms_per_step"], 3),
      "classes", {k: round(v["ms_per_step"], 3) for k, v in d["kernel_classes"].items()}, "parity_ok", d.get("parity", {}).get("ok"))
PY
done
