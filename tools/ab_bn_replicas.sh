#!/bin/bash
# GPU box: replicated fp64 BN accumulators (SCN_B200_BN_REPLICAS: 1 = one copy of the 2C column sums, 8 default)
out=gpurun_out; mkdir -p $out
for g in 1 4 16 32; do
  SCN_B200_BN_REPLICAS=$g timeout 120 python tools/bn_probe.py 10 2>&1 | tee -a $out/bn_replicas_probe.log | tail -8
done
timeout 300 python -m pytest tests -m gpu -x -q -k "batchnorm or backbone or layer_graph or dense_equivalence or wide or pruned or full_size" > $out/ab_bnrep_tests.log 2>&1
echo "tests rc=$? $(tail -1 $out/ab_bnrep_tests.log)"
for i in 1 2; do for r in 1 16; do
  SCN_B200_BN_REPLICAS=$r timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/ab_bnrep_${r}_$i.log 2>&1
  python - $out/ab_bnrep_${r}_$i.log $r <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric'):
        d=json.loads(l); k=d['kernel_classes']
        print('replicas',sys.argv[2],'step %.3f e2e %.3f inline %.3f pruned %.3f bn %.3f gemm %.3f dw %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value_inline']['ms_per_step'],d['value_pruned']['ms_per_step'],k['batchnorm']['ms_per_step'],k['conv_gemm']['ms_per_step'],k['weight_grad']['ms_per_step']))
PY
done; done
