"""GPU parity of the device-side RPN pieces (csrc/rpn.cu) against oracle/rpn_oracle.py:
anchors bit-exact (every float operation in the reference's order), per-sample scopes exact, the head's logits /
regressions / gradients within the fp32 bound of the torch CPU evaluation."""
import numpy as np
import pytest
import torch

import rpn_oracle as RO

pytestmark = pytest.mark.gpu


def _levels(scn, batch=2, C=128, seed=0):
    """two RPN levels of a small cloud: a sparse map and its 2x coarser one"""
    rng = np.random.RandomState(seed)
    c = np.concatenate([np.concatenate([(rng.rand(2500, 3) * np.array([60, 44, 14])).astype(np.int64),
                                        np.full((2500, 1), b)], 1) for b in range(batch)])
    torch.manual_seed(seed)
    t0 = scn.InputLayer(3, [64, 48, 16], 4)([torch.from_numpy(c), torch.randn(len(c), C).cuda()])
    t1 = scn.Convolution(3, C, C, 2, 2, False).cuda()(t0)
    return [t0, t1]


def test_grid_anchors_and_scopes_bit_exact():
    import sparseconvnet as scn
    levels = _levels(scn, batch=3)
    yaws = [0, -1.57, -0.785, 0.785]
    cell = [RO.generate_anchors_3d(s, yaws, None, True) for s in ([0.5, 0.3, 2.5], [1.2, 0.7, 2.9])]
    strides = [[1.0, 1.0, 1.0], [2.0, 2.0, 2.0]]
    anchors, scopes = scn.grid_anchors(levels, cell, 50, strides)
    locs = [t.get_spatial_locations() for t in levels]
    want = RO.grid_anchors(locs, cell, 50, strides)
    for a, w, s, loc in zip(anchors, want, scopes, locs):
        assert a.is_cuda and torch.equal(a.cpu(), w)
        assert torch.equal(s.cpu(), RO.examples_bidx_2_sizes(loc[:, -1], 3) * len(yaws))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp32_ffma", 1e-4), ("tf32", 3e-2)])   # tf32: measured 1.2e-2 (relative L2, ReLU-mask flips included), stated x3
@pytest.mark.parametrize("A,S", [(2, 1), (4, 2)])
def test_rpn_head_matches_torch(precision, tol, A, S):
    import sparseconvnet as scn
    scn.set_conv_precision(precision)
    try:
        levels = _levels(scn)
        head = scn.RPNHead(128, A, S).cuda()
        for p in head.parameters():
            torch.nn.init.normal_(p, std=0.05)
        assert sorted(head.state_dict()) == ["bbox_pred.bias", "bbox_pred.weight", "cls_logits.bias", "cls_logits.weight",
                                             "conv.bias", "conv.weight"]
        xs = [t.features.detach().clone().requires_grad_(True) for t in levels]
        logits, regs = head(xs)
        loss = sum((l * torch.linspace(0.5, 1.5, l.numel(), device="cuda").view_as(l)).sum() for l in logits) + \
            sum((r ** 2).sum() for r in regs)
        loss.backward()
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in head.state_dict().items()}
        cx = [x.detach().cpu().clone().requires_grad_(True) for x in xs]
        want = [RO.rpn_head(x, sd["conv.weight"], sd["conv.bias"], sd["cls_logits.weight"], sd["cls_logits.bias"],
                            sd["bbox_pred.weight"], sd["bbox_pred.bias"], A, S) for x in cx]
        closs = sum((l * torch.linspace(0.5, 1.5, l.numel()).view_as(l)).sum() for l, _ in want) + \
            sum((r ** 2).sum() for _, r in want)
        closs.backward()

        def rel(a, b):
            a, b = a.detach().cpu().double(), b.detach().double()
            if precision == "tf32":     # a ReLU mask that flips on a near-zero hidden value moves single elements by
                return float((a - b).norm() / b.norm())      # O(1): reduced precision is judged on the relative L2
            return float((a - b).abs().max() / b.abs().max())
        for (wl, wr), l, r in zip(want, logits, regs):
            assert l.shape == wl.shape and r.shape == wr.shape
            assert rel(l, wl) <= tol and rel(r, wr) <= tol
        for x, c in zip(xs, cx):
            assert rel(x.grad, c.grad) <= tol
        for k, p in head.named_parameters():
            assert rel(p.grad, sd[k].grad) <= tol, k
        # the reference's own input layout [1, C, n, 1] gives the same result
        l4, r4 = head([x.detach().t().unsqueeze(0).unsqueeze(3) for x in xs])
        assert all(torch.equal(a, b) for a, b in zip(l4, logits))
    finally:
        scn.set_conv_precision("fp32")
