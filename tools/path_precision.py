"""precision of every convolution PATH of the backbone against a float64 evaluation on the library's own rulebooks
(developer tool): submanifold 3^3 / 1^3, strided 2^3/2 convolution, deconvolution, z-collapse, at small and large
row counts, fp32 (3xTF32) beside fp32_ffma.  max|err| / max|truth| and relative L2 of y, dX, dW."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("automatic-as-built-reconstruction_b200", "oracle", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sparseconvnet as scn  # noqa: E402
import bench  # noqa: E402


def err(a, b):
    a, b = a.detach().cpu().double(), b.cpu()
    return "%.1e/%.1e" % (float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm()))


def truth(x, w, dy, rules, n_out, swap):
    x, w, dy = x.double().cuda(), w.double().cuda(), dy.double().cuda()
    y = torch.zeros(n_out, w.shape[3], dtype=torch.float64, device="cuda")
    dx, dw = torch.zeros_like(x), torch.zeros_like(w)
    for k, r in enumerate(rules):
        if len(r) == 0:
            continue
        r = r.long().cuda()
        i, o = (r[:, 1], r[:, 0]) if swap else (r[:, 0], r[:, 1])
        y.index_add_(0, o, x[i] @ w[k, 0])
        dw[k, 0] = x[i].t() @ dy[o]
        dx.index_add_(0, i, dy[o] @ w[k, 0].t())
    return y.cpu(), dx.cpu(), dw.cpu()


def run(name, locs, ss, cin, cout, make, rules_of, swap=False, pre=None):
    res = {}
    for prec in ("fp32_ffma", "fp32"):
        scn.set_conv_precision(prec)
        torch.manual_seed(0)
        x = scn.InputLayer(3, ss, 4)([locs, torch.randn(len(locs), pre[1] if pre else cin).cuda().abs()])
        if pre:
            with torch.no_grad():
                x = pre[0](x)
        x.features = x.features.detach().clone().requires_grad_(True)
        conv = make().cuda()
        y = conv(x)
        dy = torch.randn_like(y.features)
        y.features.backward(dy)
        if prec == "fp32_ffma":
            rules = rules_of(x.metadata)
            t = truth(x.features.detach(), conv.weight.detach(), dy, rules, y.features.shape[0], swap)
        res[prec] = (err(y.features, t[0]), err(x.features.grad, t[1]), err(conv.weight.grad, t[2]))
    for prec, r in res.items():
        print("%-22s rows %7d %3d->%3d %-9s y %s  dX %s  dW %s" % (name, x.features.shape[0], cin, cout, prec, *r))


g = np.load(os.path.join(ROOT, "tests", "golden", "wide_net.npz"))
small = torch.from_numpy(g["locs"].astype(np.int64))
big, _ = bench.make_batch(300000, 1, 1, 0)
for label, locs, ss in (("small", small, [512] * 3), ("big", big, [4096, 4096, 512])):
    ss1 = [s // 2 for s in ss]
    for c in (32, 64, 128):
        run(label + " subm3", locs, ss, c, c, lambda: scn.SubmanifoldConvolution(3, c, c, 3, False),
            lambda m: m.getSubmanifoldRuleBook(ss, [3, 3, 3]))
    run(label + " subm3 9->32", locs, ss, 9, 32, lambda: scn.SubmanifoldConvolution(3, 9, 32, 3, False),
        lambda m: m.getSubmanifoldRuleBook(ss, [3, 3, 3]))
    run(label + " subm1", locs, ss, 64, 128, lambda: scn.SubmanifoldConvolution(3, 64, 128, 1, False),
        lambda m: m.getSubmanifoldRuleBook(ss, [1, 1, 1]))
    run(label + " conv2/2", locs, ss, 32, 64, lambda: scn.Convolution(3, 32, 64, 2, 2, False),
        lambda m: m.getRuleBook(ss, ss1, [2, 2, 2], [2, 2, 2]))
    run(label + " conv2/2", locs, ss, 128, 128, lambda: scn.Convolution(3, 128, 128, 2, 2, False),
        lambda m: m.getRuleBook(ss, ss1, [2, 2, 2], [2, 2, 2]))
    down = scn.Convolution(3, 16, 128, 2, 2, False).cuda()
    run(label + " deconv2/2", locs, ss, 128, 128, lambda: scn.Deconvolution(3, 128, 128, 2, 2, False),
        lambda m: m.getRuleBook(ss, ss1, [2, 2, 2], [2, 2, 2]), swap=True, pre=(down, 16))
# z-collapse at the rpn map sizes: take scale 4 of the big building
locs4 = torch.cat([big[:, :3] // 16, big[:, 3:]], 1)
run("zcollapse Z=32", locs4, [256, 256, 32], 128, 128, lambda: scn.Convolution(3, 128, 128, [1, 1, 32], [1, 1, 1], False),
    lambda m: m.getRuleBook([256, 256, 32], [256, 256, 1], [1, 1, 32], [1, 1, 1]))
locs7 = torch.cat([big[:, :3] // 128, big[:, 3:]], 1)
run("tiny subm3 (112 rows)", locs7, [32, 32, 4], 256, 256, lambda: scn.SubmanifoldConvolution(3, 256, 256, 3, False),
    lambda m: m.getSubmanifoldRuleBook([32, 32, 4], [3, 3, 3]))
run("tiny conv2/2", locs7, [32, 32, 4], 256, 256, lambda: scn.Convolution(3, 256, 256, 2, 2, False),
    lambda m: m.getRuleBook([32, 32, 4], [16, 16, 2], [2, 2, 2], [2, 2, 2]))
