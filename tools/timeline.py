"""Developer tool (GPU box): CUPTI timeline of steady-state backbone steps (with the prefetcher):
GPU busy time per step, idle gaps, per-kernel totals in real (warm-cache, overlapped) conditions."""
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

scn.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
bucket = scn.GradBucket(net.parameters(), module=net)
locs, feats = bench.make_batch(300000, 1, 1, 0)
ld, fd = locs.to(dev), feats.to(dev)
pf = scn.InputPrefetcher(net.prepare)


def steps(n):
    pf.submit(ld)
    for i in range(n):
        p = pf.get()
        if i + 1 < n:
            pf.submit(ld)
        bucket.zero()
        rpn, roi = net([p, fd])
        loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
        loss.backward()


steps(4)
torch.cuda.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    steps(N)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
span = iv[-1][1] - iv[0][0]
print("steps %d: span %.2f ms/step, GPU busy (union) %.2f ms/step, kernels %d/step" %
      (N, span / N / 1e3, busy / N / 1e3, len(ev) // N))
tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][:70]
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
S = sum(v[1] for v in tot.values())
print("sum of kernel durations %.2f ms/step" % (S / N / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:26]:
    print("%-70s %5d %9.1f us/step %5.1f%%" % (k, v[0] // N, v[1] / N, 100 * v[1] / S))
pf.close()

# ---- idle gaps of the union timeline and busy time per stream -----------------------------------
named = sorted((e.time_range.start, e.time_range.end, e.name.split("(")[0][:48],
                getattr(e, "stream", None) if hasattr(e, "stream") else None) for e in ev)
gaps, cur_e, last = [], None, None
for s, e, nm, st in named:
    if cur_e is not None and s > cur_e:
        gaps.append((s - cur_e, last, nm, cur_e - named[0][0]))
    if cur_e is None or e > cur_e:
        cur_e, last = e, nm
print("idle gaps: %d, total %.2f ms/step" % (len(gaps), sum(g[0] for g in gaps) / N / 1e3))
for g in sorted(gaps, key=lambda g: -g[0])[:24]:
    print("  %7.1f us at t=%8.1f us   after %-48s before %s" % (g[0], g[3], g[1], g[2]))
hist = collections.Counter(min(int(g[0] // 5) * 5, 100) for g in gaps)
print("gap histogram (us bucket: count/step):", {k: v // N for k, v in sorted(hist.items())})
