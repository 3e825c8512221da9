"""Developer tool (GPU box): CUPTI kernel totals of the integer work of one batch (FPN_Net.prepare: input grid, 12
coarser grids, 25 rulebooks with their tile books) run alone."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
PTS = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
FLOORS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
locs, feats = bench.make_batch(PTS, FLOORS, 1, 0)
ld = locs.to(dev)
for _ in range(3):
    net.prepare(ld)
torch.cuda.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        net.prepare(ld)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][:60]
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
S = sum(v[1] for v in tot.values())
span = max(e.time_range.end for e in ev) - min(e.time_range.start for e in ev)
print("build: %d kernels / copies, %.2f ms of kernel time, span %.2f ms per batch" % (len(ev) // N, S / N / 1e3, span / N / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-60s %5d %8.1f us %5.1f%%  (%.1f us each)" % (k, v[0] // N, v[1] / N, 100 * v[1] / S, v[1] / v[0]))
