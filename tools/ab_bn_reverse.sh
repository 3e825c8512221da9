#!/bin/bash
# NOTE: the knob this script flips exists only with profiles/experiments/r02_bn_reverse_traversal.patch.txt applied (the experiment was
# measured and not kept - profiles/experiments/README.md); on the shipped tree both arms run the same code.
# GPU box: A/B of the BN traversal order (SCN_B200_BN_REVERSE=0: statistics and apply both front to back; 1: the
# statistics pass walks the rows back to front), two interleaved runs each; the BN parity tests run on the default.
out=gpurun_out; mkdir -p $out
timeout 200 python -m pytest tests -m gpu -x -q -k "batchnorm or backbone or layer_graph or dense_equivalence" > $out/ab_bnrev_tests.log 2>&1
echo "tests rc=$? $(tail -1 $out/ab_bnrev_tests.log)"
for i in 1 2; do for r in 0 1; do
  SCN_B200_BN_REVERSE=$r timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/ab_bnrev_${r}_$i.log 2>&1
  python - $out/ab_bnrev_${r}_$i.log $r <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric'):
        d=json.loads(l); k=d['kernel_classes']
        print('rev',sys.argv[2],'step %.3f e2e %.3f inline %.3f bn %.3f gemm %.3f dw %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value_inline']['ms_per_step'],k['batchnorm']['ms_per_step'],k['conv_gemm']['ms_per_step'],k['weight_grad']['ms_per_step']))
PY
done; done
