"""Developer tool (GPU box): is a backbone step bit-reproducible?  Runs the wide golden net forward + backward several
times per precision mode and compares every output map and parameter gradient bit for bit with the first run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scn_oracle as O  # noqa: E402
import sparseconvnet as scn  # noqa: E402
from test_full_parity import WIDE_CFG, _fpn  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "wide_net.npz"))
sd = O.seeded_state_dict({k[6:]: g[k] for k in g.files if k.startswith("shape/")}, int(g["seed"]))
locs, feats = torch.from_numpy(g["locs"].astype(np.int64)), torch.from_numpy(g["feats"]).cuda()
for prec in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "fp32_ffma", "tf32", "bf16"]):
    scn.set_conv_precision(prec)
    net = _fpn(scn, WIDE_CFG, 32)
    net.load_state_dict(sd)
    net = net.cuda().train()
    first = None
    for rep in range(4):
        net.zero_grad(set_to_none=True)
        rpn, roi = net([locs, feats])
        sum((m.features ** 2).sum() for m in list(rpn) + list(roi)).backward()
        torch.cuda.synchronize()
        cur = [m.features.detach().clone() for m in list(rpn) + list(roi)] + \
              [p.grad.detach().clone() for _, p in net.named_parameters() if p.grad is not None]
        if first is None:
            first = cur
        else:
            bad = [i for i, (a, b) in enumerate(zip(first, cur)) if not torch.equal(a, b)]
            print("%s run %d: %d of %d tensors differ from run 0%s" % (
                prec, rep, len(bad), len(cur),
                "" if not bad else "  (first: #%d, max |diff| %.3g of max %.3g)" % (
                    bad[0], float((first[bad[0]] - cur[bad[0]]).abs().max()), float(first[bad[0]].abs().max()))))
