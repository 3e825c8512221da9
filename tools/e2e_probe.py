"""Developer tool (GPU box): why is the host-input (e2e) step slower than the device-input step?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

scn.set_conv_precision("tf32")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
bucket = scn.GradBucket(net.parameters())
locs, feats = bench.make_batch(300000, 1, 1, 0)
lp, fp = locs.pin_memory(), feats.pin_memory()
ld, fd = locs.to(dev), feats.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(host, item, use_bucket, n=6):
    t = [0.0] * 5
    for it in range(n + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        flush.fill_(1)
        if use_bucket:
            bucket.zero()
        else:
            net.zero_grad(set_to_none=True)
        c, f = (lp, fp.to(dev, non_blocking=True)) if host else (ld, fd)
        t1 = time.perf_counter()
        rpn, roi = net([c, f])
        loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        if item:
            loss.item()
        t4 = time.perf_counter()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        if it >= 2:
            for i, (a, b) in enumerate(((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))):
                t[i] += (b - a) / n * 1e3
    print("host=%d item=%d bucket=%d : prep %.2f  fwd %.2f  bwd(launch) %.2f  item %.2f  final sync %.2f  | total %.2f ms"
          % (host, item, use_bucket, t[0], t[1], t[2], t[3], t[4], sum(t)))


for cfg in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (1, 1, 1), (0, 1, 1)):
    run(*cfg)
