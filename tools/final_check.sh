#!/bin/bash
# GPU box: the round's closing pass - GPU tests, smoke, the default bench line, then (each only after the plain
# command exited 0) the ncu launch list of the timed steps, one --set full capture of the two tensor-core kernels on
# the 128 -> 128 scale-1 layer and the DRAM bytes of one step's gather-GEMM launches.  Numbers printed under ncu are
# never bench values.   gpurun --timeout 900 -- 'bash tools/final_check.sh TAG'
tag=${1:-r02_final}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_gputests.log
timeout 120 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
timeout 300 python bench.py > $out/${tag}_bench.log 2>&1; rc=$?; echo "bench rc=$rc"
[ $rc -ne 0 ] && exit $rc
SCN_BENCH_CUPROF=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
timeout 120 python tools/gemm_probe.py 128 1 fp32 1 > $out/${tag}_probe.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_osgemm_tf32|k_dw_tf32' -s 6 -c 3 -f \
  -o $out/${tag}_gemm_probe128 python tools/gemm_probe.py 128 1 fp32 1 > $out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i $out/${tag}_gemm_probe128.ncu-rep --page raw --csv > $out/${tag}_ncu_full_gemm_probe128.csv 2>/dev/null
SCN_BENCH_CUPROF=1 timeout 300 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:k_osgemm --csv --log-file $out/${tag}_gemm_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
  > $out/${tag}_ncu_dram.log 2>&1
echo "ncu dram rc=$?"
tail -1 $out/${tag}_gputests.log; tail -2 $out/${tag}_smoke.log; grep -o '"ms_per_step": [0-9.]*' $out/${tag}_bench.log | head -3
