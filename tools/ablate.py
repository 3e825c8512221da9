"""Developer tool (GPU box): what does each part of the step cost on the critical path?  Times the
backbone step (CUDA events, 10 steps) with the prefetcher, with ONE PreparedInput reused (no integer
work at all), forward only, and with the build run inline."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

scn.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
bucket = scn.GradBucket(net.parameters(), module=net)
locs, feats = bench.make_batch(int(os.environ.get("POINTS", 300000)), 1, int(os.environ.get("BATCH", 1)), 0)
ld, fd = locs.to(dev), feats.to(dev)
pf = scn.InputPrefetcher(net.prepare)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fwd_bwd(p, backward=True):
    bucket.zero()
    rpn, roi = net([p, fd])
    if backward:
        loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
        loss.backward()


def timed(fn, n=10):
    fn(3)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn(n)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def with_prefetch(n):
    pf.submit(ld)
    for i in range(n):
        flush.fill_(1)
        p = pf.get()
        if i + 1 < n:
            pf.submit(ld)
        fwd_bwd(p)


def with_prefetch2(n):
    pf.submit(ld)
    if n > 1:
        pf.submit(ld)
    for i in range(n):
        flush.fill_(1)
        p = pf.get()
        if i + 2 < n:
            pf.submit(ld)
        fwd_bwd(p)


P = net.prepare(ld)


def reuse(n):
    for i in range(n):
        flush.fill_(1)
        fwd_bwd(P)


def reuse_fwd(n):
    for i in range(n):
        flush.fill_(1)
        with torch.no_grad():
            fwd_bwd(P, False)


def inline(n):
    for i in range(n):
        flush.fill_(1)
        fwd_bwd(ld)


def build_only(n):
    for i in range(n):
        net.prepare(ld)


for name, fn in (("prefetch", with_prefetch), ("prefetch, 2 batches ahead", with_prefetch2), ("reuse prepared (no integer work)", reuse),
                 ("reuse, forward only (no_grad)", reuse_fwd), ("inline build", inline), ("build only", build_only)):
    print("%-36s %7.2f ms/step" % (name, timed(fn)))
pf.close()
