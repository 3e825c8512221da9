#!/bin/bash
# Developer tool (GPU box): the gather-GEMM's L2 prefetch distance (steps ahead; 0 = off) on single layers
for pf in 0 6 12 24; do
  for c in 128 64 32; do
    echo "PF=$pf $(SCN_B200_GEMM_PF=$pf python tools/gemm_probe.py $c 1 ${1:-fp32,fp32_split,tf32} 5 2>&1 | grep 'scale 1' | sed 's/nActive.*ratio 1.25)//' | tr '\n' '|')"
  done
done
