"""Developer tool (GPU box): wall-clock breakdown of one backbone step - forward / backward, device vs
pinned-host inputs - and per-layer tile-book efficiency.  Not part of the product or the tests."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    scn.set_conv_precision(prec)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128,
                      residual_blocks=True, fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                      downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                      rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
    locs, feats = bench.make_batch(300000, 1, 1, 0)
    inputs = {"dev": (locs.to(dev), feats.to(dev)), "host": (locs.pin_memory(), feats.pin_memory())}

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for mode in ("dev", "host", "dev"):
        c, f = inputs[mode]
        acc = [0.0] * 4
        n = 6
        for it in range(n + 2):
            net.zero_grad(set_to_none=True)
            t0 = sync()
            ff = f.to(dev, non_blocking=True)
            rpn, roi = net([c, ff])
            loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
            t1c = time.perf_counter()
            t1 = sync()
            loss.backward()
            t2c = time.perf_counter()
            t2 = sync()
            if it >= 2:
                acc[0] += t1 - t0
                acc[1] += t2 - t1
                acc[2] += t1c - t0
                acc[3] += t2c - t1
        print("%-5s %s fwd %.2f ms (cpu-side %.2f)  bwd %.2f ms (cpu-side launch %.2f)" %
              (mode, prec, acc[0] / n * 1e3, acc[2] / n * 1e3, acc[1] / n * 1e3, acc[3] / n * 1e3))

    # tile-book efficiency of the last metadata
    m = rpn[0].metadata
    ss = list(bench.FULL_SCALE)
    print("scale  nActive  pairs  tile_rows(fwd)  ratio")
    for s in range(9):
        st = m.ruleBookStats(0, ss, [3, 3, 3])
        print("subm3 s%d %8d %9d %9d  %.2f" % (s, m.getNActive(ss), st["pairs"], st["tile_rows"],
                                                st["tile_rows"] / max(st["pairs"], 1)))
        if s < 8:
            st = m.ruleBookStats(1, ss, [2, 2, 2], [2, 2, 2])
            print("conv2 s%d          %9d %9d  %.2f" % (s, st["pairs"], st["tile_rows"], st["tile_rows"] / max(st["pairs"], 1)))
            ss = [v // 2 for v in ss]


if __name__ == "__main__":
    main()
