"""Developer tool (GPU box): one submanifold 3^3 convolution C->C on the scale-s grid of the benchmark
building (points quantised at 50/2^s voxels per metre), forward + backward, timed per precision with
CUDA events.  Used for ncu captures of a single big gather-GEMM / weight-gradient launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
s = int(sys.argv[2]) if len(sys.argv) > 2 else 1
precs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "tf32"]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
xyz = bench.building(300000)
a = xyz * (50 / 2 ** s)
a -= a.min(0)
locs = torch.from_numpy(a).long()
locs = torch.cat([locs, torch.zeros(len(locs), 1, dtype=torch.long)], 1)
ss = [4096 >> s, 4096 >> s, 512 >> s]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
feats = torch.randn(len(locs), C, device=dev)
conv = scn.SubmanifoldConvolution(3, C, C, 3, False).to(dev)
for prec in precs:
    scn.set_conv_precision(prec)
    x = scn.InputLayer(3, ss, 4)([locs, feats])
    x.features.requires_grad_(True)
    y = conv(x)                                  # builds the rulebook
    g = torch.ones_like(y.features)
    tf, tb = [], []
    for it in range(reps + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        e[0].record()
        y = conv(x)
        e[1].record()
        torch.cuda.synchronize()
        e[2].record()
        y.features.backward(g)
        e[3].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[2].elapsed_time(e[3]))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            y = conv(x)
            y.features.backward(g)
        torch.cuda.synchronize()
    kt = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            kt.setdefault(ev.name.split("(")[0][-40:], []).append(ev.time_range.end - ev.time_range.start)
    print("   kernels (us): " + "  ".join("%s x%d %.1f" % (k, len(v) // 3, np.median(v)) for k, v in kt.items()))
    st = x.metadata.ruleBookStats(0, ss, [3, 3, 3])
    f, b = np.median(tf[2:]), np.median(tb[2:])
    steps = st["tile_rows"] / 128 * (C // 32)
    print("%s C=%d scale %d: nActive %d pairs %d tile_rows %d (ratio %.2f)  fwd %.3f ms (%.2f us/step/SM)  bwd(dX+dW) %.3f ms"
          % (prec, C, s, x.features.shape[0], st["pairs"], st["tile_rows"], st["tile_rows"] / st["pairs"], f,
             f * 1e3 / (steps / 148), b))
