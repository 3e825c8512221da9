"""Developer tool (GPU box): where does the main thread's enqueue time go?  Wraps every C-ABI entry point
with a perf_counter pair and reports host seconds per step inside the library (launch API included) vs
the Python around it."""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402
from sparseconvnet import _lib  # noqa: E402

scn.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
acc = collections.defaultdict(lambda: [0, 0.0])


class Timed(object):
    def __init__(self, lib):
        object.__setattr__(self, "_l", lib)
        object.__setattr__(self, "_c", {})

    def __getattr__(self, name):
        c = self._c.get(name)
        if c is None:
            f = getattr(self._l, name)
            slot = acc[name]

            def c(*a):
                t = time.perf_counter()
                r = f(*a)
                slot[1] += time.perf_counter() - t
                slot[0] += 1
                return r
            self._c[name] = c
        return c


timed = Timed(_lib.lib)
_lib.lib = timed
import sparseconvnet.SCN as S  # noqa: E402
S.lib = timed
dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
bucket = scn.GradBucket(net.parameters())
locs, feats = bench.make_batch(300000, 1, 1, 0)
ld, fd = locs.to(dev), feats.to(dev)
pf = scn.InputPrefetcher(net.prepare)


def run(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pf.submit(ld)
    for i in range(n):
        p = pf.get()
        if i + 1 < n:
            pf.submit(ld)
        bucket.zero()
        rpn, roi = net([p, fd])
        loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
        loss.backward()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n, (time.perf_counter() - t0) / n


run(5)
for v in acc.values():
    v[0], v[1] = 0, 0.0
N = 20
enq, wall = run(N)
print("enqueue %.2f ms/step, wall %.2f ms/step" % (enq * 1e3, wall * 1e3))
tot = 0.0
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-36s %5.1f calls/step %8.1f us/step %6.1f us/call" % (k, c / N, t / N * 1e6, t / c * 1e6))
    if k != "scn_build_plan":
        tot += t
print("inside the library on the main thread: %.2f ms/step" % (tot / N * 1e3))
pf.close()
