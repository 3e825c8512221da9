#!/bin/bash
# NOTE: the knob this script flips exists only with profiles/experiments/r02_fused_residual_add_v2.patch.txt applied (the experiment was
# measured and not kept - profiles/experiments/README.md); on the shipped tree both arms run the same code.
# GPU box: the fused conv -> add epilogue (SCN_B200_FUSE_ADD=1, default) against separate k_add launches (=0):
# bit-identity / parity tests first, then two interleaved bench runs each.
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "fused_adds or layer_graph or backbone or wide or dense_equivalence or pruned or grad_sink" > $out/ab_fuse_tests.log 2>&1
rc=$?; echo "tests rc=$rc $(tail -1 $out/ab_fuse_tests.log)"
[ $rc -ne 0 ] && { tail -40 $out/ab_fuse_tests.log; exit $rc; }
for i in 1 2; do for r in 0 1; do
  SCN_B200_FUSE_ADD=$r timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/ab_fuse_${r}_$i.log 2>&1
  python - $out/ab_fuse_${r}_$i.log $r <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric'):
        d=json.loads(l); k=d['kernel_classes']
        print('fuse',sys.argv[2],'step %.3f e2e %.3f inline %.3f pruned %.3f launches %d bn %.3f gemm %.3f dw %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value_inline']['ms_per_step'],d['value_pruned']['ms_per_step'],d['gpu_launches'],k['batchnorm']['ms_per_step'],k['conv_gemm']['ms_per_step'],k['weight_grad']['ms_per_step']))
PY
done; done
