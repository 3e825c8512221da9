"""Developer tool (GPU box): BN(+ReLU) forward / backward on the large BN shapes of the benchmark building, each kernel
pair timed with CUDA events after an L2 flush, for the current SCN_B200_BN_STATS_* knobs.  Prints GB/s against the
SURVEY 8d algorithmic bytes (3 n C s forward, 5 n C s backward)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import sparseconvnet as scn  # noqa: E402
from sparseconvnet._lib import lib, ptr, stream, check  # noqa: E402

SHAPES = [(278639, 32), (227288, 64), (123489, 64), (38672, 128), (9544, 128), (278639, 128), (227288, 128)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = "replicas=%s" % os.environ.get("SCN_B200_BN_REPLICAS", "8")
tot_f = tot_b = 0.0
for n, C in SHAPES:
    x = torch.randn(n, C, device=dev)
    y = torch.empty_like(x)
    dy = torch.randn(n, C, device=dev)
    dx = torch.empty_like(x)
    sm, si = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    w, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    dw, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    tf = tb = 0.0
    for it in range(reps + 2):
        flush.fill_(1)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        check(lib.scn_batchnorm_forward(ptr(x), ptr(y), ptr(sm), ptr(si), ptr(rm), ptr(rv), ptr(w), ptr(b),
                                        ctypes.c_float(1e-4), ctypes.c_float(0.95), 1, ctypes.c_float(0.0), n, C, stream()))
        e[1].record()
        flush.fill_(1)
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record()
        check(lib.scn_batchnorm_backward_fused(ptr(x), ptr(dx), None, ptr(dy), ptr(sm), ptr(si), ptr(w), ptr(b), 1,
                                               ptr(dw), ptr(db), ctypes.c_float(0.0), n, C, None, stream()))
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf += e[0].elapsed_time(e[1])
            tb += e2.elapsed_time(e[2])
    tf, tb = tf / reps * 1e3, tb / reps * 1e3
    tot_f += tf
    tot_b += tb
    print("%s  n=%6d C=%3d  fwd %6.1f us %5.2f TB/s   bwd %6.1f us %5.2f TB/s"
          % (tag, n, C, tf, 3 * n * C * 4 / tf / 1e6, tb, 5 * n * C * 4 / tb / 1e6))
print("%s  total fwd %.1f us bwd %.1f us" % (tag, tot_f, tot_b))
