"""per-layer precision of the convolution kernels against a float64 evaluation (developer tool):
    python tools/layer_precision.py  ->  max|err|/max|truth| of y, dX, dW per precision mode and shape"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("automatic-as-built-reconstruction_b200", "oracle", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scn_oracle as O  # noqa: E402
import sparseconvnet as scn  # noqa: E402


def rel(a, b):
    return float((a.detach().cpu().double() - b.cpu()).abs().max() / b.abs().max())


def truth(x, w, dy, rules, n_out, swap=False):
    x, w, dy = x.double().cuda(), w.double().cuda(), dy.double().cuda()
    y = torch.zeros(n_out, w.shape[3], dtype=torch.float64, device="cuda")
    dx, dw = torch.zeros_like(x), torch.zeros_like(w)
    for k, r in enumerate(rules):
        if len(r) == 0:
            continue
        r = torch.as_tensor(np.asarray(r), dtype=torch.int64).cuda()
        i, o = (r[:, 1], r[:, 0]) if swap else (r[:, 0], r[:, 1])
        y.index_add_(0, o, x[i] @ w[k, 0])
        dw[k, 0] = x[i].t() @ dy[o]
        dx.index_add_(0, i, dy[o] @ w[k, 0].t())
    return y.cpu(), dx.cpu(), dw.cpu()


g = np.load(os.path.join(ROOT, "tests", "golden", "wide_net.npz"))
locs = torch.from_numpy(g["locs"].astype(np.int64))
for n_pts, ss in ((len(locs), [512] * 3), (0, None)):
    if n_pts == 0:
        import bench
        locs, _ = bench.make_batch(300000, 1, 1, 0)
        locs = torch.cat([locs[:, :3] // 2, locs[:, 3:]], 1)
        ss = [2048, 2048, 256]
    for cin, cout in ((64, 64), (128, 128), (32, 64)):
        for prec in ("fp32_ffma", "fp32", "tf32", "bf16"):
            scn.set_conv_precision(prec)
            torch.manual_seed(0)
            x = scn.InputLayer(3, ss, 4)([locs, torch.randn(len(locs), cin).cuda()])
            x.features.requires_grad_(True)
            conv = scn.SubmanifoldConvolution(3, cin, cout, 3, False).cuda()
            y = conv(x)
            dy = torch.randn_like(y.features)
            y.features.backward(dy)
            loc = x.get_spatial_locations().numpy()
            if prec == "fp32_ffma":
                rules = O.submanifold_rules(loc, ss, [3] * 3)
                ty, tdx, tdw = truth(x.features.detach(), conv.weight.detach(), dy, rules, len(loc))
            print("rows %7d  %3d->%3d  %-9s  y %.2e  dX %.2e  dW %.2e" %
                  (len(loc), cin, cout, prec, rel(y.features, ty), rel(x.features.grad, tdx), rel(conv.weight.grad, tdw)))
