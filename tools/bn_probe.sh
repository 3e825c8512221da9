#!/bin/bash
# NOTE: SCN_B200_BN_STATS_CAP / _UNROLL were knobs of an experimental build of bn.cu (statistics blocks per SM, row slots
# unrolled; results in profiles/experiments/r02_bn_probe_grid_unroll.txt: the shipped 4 / 4 is the best pair) - the shipped
# library ignores them; tools/bn_probe.py itself and SCN_B200_BN_REPLICAS (tools/ab_bn_replicas*.sh) work on the shipped tree.
out=gpurun_out; mkdir -p $out
for cfg in "4 4" "6 4" "8 4" "4 8" "8 8" "2 4"; do set -- $cfg
  SCN_B200_BN_STATS_CAP=$1 SCN_B200_BN_STATS_UNROLL=$2 timeout 120 python tools/bn_probe.py 10 2>&1 | tee -a $out/bn_probe.log | tail -8
done
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:k_bn_stats_vec' -s 4 -c 4 -f -o $out/bn_stats_probe python tools/bn_probe.py 1 > $out/bn_probe_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i $out/bn_stats_probe.ncu-rep --page raw --csv > $out/bn_stats_probe_raw.csv 2>/dev/null
ncu -i $out/bn_stats_probe.ncu-rep --page details 2>/dev/null | grep -i -A3 "stall\|Issue Slot\|Warp Cycles Per\|Est. Speedup\|Achieved Occupancy\|DRAM Throughput\|Duration" | head -150 > $out/bn_stats_probe_details.txt
