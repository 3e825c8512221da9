"""The oracle against plain dense PyTorch ops: on a fully active grid - and, with a mask, on a
sparse one - the sparse layers of the reference equal conv3d / conv_transpose3d / batch_norm.  A reference-free pin of
the oracle's orientation, weight layout (K enumerates the filter box with z fastest, W[k] is [Cin, Cout]) and BN
arithmetic; the compiled reference itself satisfied the same identities during the survey.  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import scn_oracle as O

TOL = 1e-5


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def _grid(S, B, occupancy, seed):
    rng = np.random.RandomState(seed)
    g = np.stack(np.meshgrid(np.arange(S), np.arange(S), np.arange(S), indexing="ij"), -1).reshape(-1, 3)
    parts = []
    for b in range(B):
        keep = rng.rand(len(g)) < occupancy
        c = g[keep][rng.permutation(int(keep.sum()))]          # any point order inside a sample
        parts.append(np.concatenate([c, np.full((len(c), 1), b)], 1))
    return np.concatenate(parts).astype(np.int64)


@pytest.mark.parametrize("occupancy", [1.0, 0.15])
def test_convolutions_equal_dense_ops(occupancy):
    S, B, cin, cout = 8 if occupancy == 1.0 else 16, 2, 5, 7
    torch.manual_seed(0)
    c = _grid(S, B, occupancy, 1)
    x = torch.randn(len(c), cin)
    dense = O.sparse_to_dense(x, c, [S] * 3, B)
    mask = O.sparse_to_dense(torch.ones(len(c), 1), c, [S] * 3, B)

    # submanifold 3^3: cross-correlation with zero padding, evaluated at the active sites only
    w = torch.randn(27, 1, cin, cout)
    y = O.conv_forward(x, w, O.submanifold_rules(c, [S] * 3, [3] * 3), len(c))
    want = mask * F.conv3d(dense, w.view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2), padding=1)
    assert _rel(O.sparse_to_dense(y, c, [S] * 3, B), want) <= TOL

    # strided 2^3 / 2: plain strided conv3d, and the new sites are exactly the non-empty 2^3 cells
    w2 = torch.randn(8, 1, cin, cout)
    oc, rules = O.conv_rules(c, [S] * 3, [2] * 3, [2] * 3, [S // 2] * 3)
    z = O.conv_forward(x, w2, rules, len(oc))
    want = F.conv3d(dense, w2.view(2, 2, 2, cin, cout).permute(4, 3, 0, 1, 2), stride=2)
    assert _rel(O.sparse_to_dense(z, oc, [S // 2] * 3, B), want) <= TOL
    assert len(oc) == int((F.max_pool3d(mask, 2) > 0).sum())
    assert (np.diff(oc[:, 3]) >= 0).all()                       # batch-contiguous ascending

    # deconvolution 2^3 / 2 back onto the fine sites: the same rulebook with the roles swapped
    w3 = torch.randn(8, 1, cout, cin)
    u = O.conv_forward(z, w3, rules, len(c), swap=True)
    want = mask * F.conv_transpose3d(O.sparse_to_dense(z, oc, [S // 2] * 3, B),
                                     w3.view(2, 2, 2, cout, cin).permute(3, 4, 0, 1, 2), stride=2)
    assert _rel(O.sparse_to_dense(u, c, [S] * 3, B), want) <= TOL

    # z-collapse [1,1,S] / 1
    w4 = torch.randn(S, 1, cin, cout)
    zc, zr = O.conv_rules(c, [S] * 3, [1, 1, S], [1, 1, 1], [S, S, 1])
    v = O.conv_forward(x, w4, zr, len(zc))
    want = F.conv3d(dense, w4.view(1, 1, S, cin, cout).permute(4, 3, 0, 1, 2))
    assert _rel(O.sparse_to_dense(v, zc, [S, S, 1], B), want) <= TOL


def test_conv_backward_is_the_adjoint():
    """<conv(x), dy> == <x, dX> and == <W, dW>: the gradients CPU/Convolution.cpp:82-115 computes are the adjoints of
    the forward map in x and in W (checked in float64 arithmetic on float32 operands)"""
    S, B, cin, cout = 10, 2, 6, 4
    torch.manual_seed(1)
    c = _grid(S, B, 0.3, 2)
    rules = O.submanifold_rules(c, [S] * 3, [3] * 3)
    x, w, dy = torch.randn(len(c), cin), torch.randn(27, 1, cin, cout), torch.randn(len(c), cout)
    y = O.conv_forward(x, w, rules, len(c))
    dx, dw, _ = O.conv_backward(x, dy, w, rules)
    lhs = float((y.double() * dy.double()).sum())
    assert abs(lhs - float((x.double() * dx.double()).sum())) <= 1e-4 * abs(lhs)
    assert abs(lhs - float((w.double() * dw.double()).sum())) <= 1e-4 * abs(lhs)
    oc, r2 = O.conv_rules(c, [S] * 3, [2] * 3, [2] * 3, [S // 2] * 3)
    w2, dz = torch.randn(8, 1, cin, cout), torch.randn(len(oc), cout)
    z = O.conv_forward(x, w2, r2, len(oc))
    dx2, dw2, _ = O.conv_backward(x, dz, w2, r2)
    lhs = float((z.double() * dz.double()).sum())
    assert abs(lhs - float((x.double() * dx2.double()).sum())) <= 1e-4 * abs(lhs)
    assert abs(lhs - float((w2.double() * dw2.double()).sum())) <= 1e-4 * abs(lhs)


@pytest.mark.parametrize("leak", [0.0, 0.333])
def test_batchnorm_equals_dense_formula(leak):
    """train: biased variance normalises, unbiased variance feeds the running buffer, momentum weighs the OLD value
    (CPU/BatchNormalization.cpp:19-40); backward against autograd of the same formula"""
    torch.manual_seed(2)
    n, C, eps, mom = 500, 12, 1e-4, 0.95
    x = torch.randn(n, C) * 3 + 1
    g, b = torch.rand(C) + 0.5, torch.randn(C)
    rm, rv = torch.randn(C), torch.rand(C) + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    y, mean, invstd = O.bn_forward(x, g, b, rm, rv, eps, mom, True, leak)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre = (xr - xr.mean(0)) / torch.sqrt(xr.var(0, unbiased=False) + eps) * gr + br
    want = torch.where(pre > 0, pre, pre * leak)
    assert _rel(y, want.detach()) <= TOL
    assert _rel(rm, mom * rm0 + (1 - mom) * x.mean(0)) <= TOL
    assert _rel(rv, mom * rv0 + (1 - mom) * x.var(0, unbiased=True)) <= TOL
    dy = torch.randn(n, C)
    want.backward(dy)
    dx, dg, db = O.bn_backward(x, y, dy, mean, invstd, g, leak)
    assert _rel(dx, xr.grad) <= 1e-4 and _rel(dg, gr.grad) <= 1e-4 and _rel(db, br.grad) <= 1e-4
    # eval with the buffers as passed (:41-46)
    ye, _, _ = O.bn_forward(x, g, b, rm.clone(), rv.clone(), eps, mom, False, leak)
    pre = (x - rm) / torch.sqrt(rv + eps) * g + b
    assert _rel(ye, torch.where(pre > 0, pre, pre * leak)) <= TOL


def test_input_layer_round_trip():
    """InputLayer(mode 4) -> SparseToDense of distinct voxels is the identity; duplicates average"""
    S = 6
    c = _grid(S, 2, 0.5, 3)
    x = torch.randn(len(c), 3)
    sites, point_row, header, table = O.input_layer_rules(np.concatenate([c, c[:7]]), mode=4)
    feats = torch.cat([x, x[:7] + 2.0])
    y = O.input_layer_forward(feats, header, table)
    assert y.shape[0] == len(c) and np.array_equal(sites, c)         # first-occurrence order = the order given
    assert np.array_equal(point_row, np.concatenate([np.arange(len(c)), np.arange(7)]))
    want = x.clone()
    want[:7] += 1.0                                                 # mean of x and x + 2
    assert _rel(y, want) <= TOL
    dense = O.sparse_to_dense(y, sites, [S] * 3, 2)
    assert _rel(dense[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]], want) <= TOL
    assert float(dense.abs().sum()) == pytest.approx(float(want.abs().sum()), rel=1e-5)
