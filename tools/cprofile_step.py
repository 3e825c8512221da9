"""Developer tool (GPU box): cProfile of the host side of backbone steps."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402

scn.set_conv_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(bench.FULL_SCALE, 3, ["xyz", "color", "normal"], 1, bench.PLANES, nPlaneM=128, residual_blocks=True,
                  fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                  downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=bench.RPN_SIZES, voxel_scale=50,
                  rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False).to(dev).train()
locs, feats = bench.make_batch(300000, 1, 1, 0)
c, f = locs.to(dev), feats.to(dev)


pf = scn.InputPrefetcher(net.prepare)
pf.submit(c)


def step():
    net.zero_grad(set_to_none=True)
    p = pf.get()
    pf.submit(c)
    rpn, roi = net([p, f])
    loss = sum((m.features ** 2).sum() for m in list(rpn) + list(roi))
    loss.backward()


for _ in range(25):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
