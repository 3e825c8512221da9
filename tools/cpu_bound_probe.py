"""Developer tool (GPU box): is the steady-state step CPU- or GPU-bound?  Times the main thread's enqueue
work per step (no syncs) against the wall time per step and the GPU-only time (CUDA events)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

scn.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
bucket = scn.GradBucket(net.parameters())
locs, feats = bench.make_batch(300000, 1, 1, 0)
ld, fd = locs.to(dev), feats.to(dev)
pf = scn.InputPrefetcher(net.prepare)


def run(n, sync_each):
    torch.cuda.synchronize()
    t_get = t_fwd = t_bwd = 0.0
    t0 = time.perf_counter()
    pf.submit(ld)
    for i in range(n):
        a = time.perf_counter()
        p = pf.get()
        if i + 1 < n:
            pf.submit(ld)
        b = time.perf_counter()
        bucket.zero()
        rpn, roi = net([p, fd])
        loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
        c = time.perf_counter()
        loss.backward()
        d = time.perf_counter()
        if sync_each:
            torch.cuda.synchronize()
        t_get += b - a
        t_fwd += c - b
        t_bwd += d - c
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("sync_each=%d: wall %.2f ms/step | main thread: get %.2f  fwd enqueue %.2f  bwd enqueue %.2f ms" %
          (sync_each, wall / n * 1e3, t_get / n * 1e3, t_fwd / n * 1e3, t_bwd / n * 1e3))


run(5, 0)
run(20, 0)
run(20, 1)
run(20, 0)
pf.close()
