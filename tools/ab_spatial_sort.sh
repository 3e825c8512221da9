#!/bin/bash
# NOTE: the knob this script flips exists only with profiles/experiments/r02_spatial_sort_scale1.patch.txt applied (the experiment was
# measured and not kept - profiles/experiments/README.md); on the shipped tree both arms run the same code.
# GPU box: Morton renumbering of the first large conv-created scale (SCN_B200_SPATIAL_SORT=1, default) against
# first-touch order (=0): parity tests on the default first, then two interleaved bench runs each.
out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -x -q > $out/ab_spatial_tests.log 2>&1
rc=$?; echo "tests rc=$rc $(tail -1 $out/ab_spatial_tests.log)"
[ $rc -ne 0 ] && { tail -40 $out/ab_spatial_tests.log; exit $rc; }
for i in 1 2; do for r in 0 1; do
  SCN_B200_SPATIAL_SORT=$r timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/ab_spatial_${r}_$i.log 2>&1
  python - $out/ab_spatial_${r}_$i.log $r <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric'):
        d=json.loads(l); k=d['kernel_classes']
        print('spatial',sys.argv[2],'step %.3f e2e %.3f inline %.3f pruned %.3f launches %d bn %.3f gemm %.3f dw %.3f rules %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value_inline']['ms_per_step'],d['value_pruned']['ms_per_step'],d['gpu_launches'],k['batchnorm']['ms_per_step'],k['conv_gemm']['ms_per_step'],k['weight_grad']['ms_per_step'],k['rulebook']['ms_per_step']))
PY
done; done
