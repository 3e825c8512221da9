"""Pin the oracle's compute restatement and the reference-op backbone driver against the compiled
reference (oracle/_ref/*.so, built from /root/reference by oracle/build_ref.py) and against the
golden outputs of the reference's real fpn_net.py (tests/golden/small_net*.npz)."""
import numpy as np
import pytest
import torch

import ref_backbone as RB
import scn_oracle as O

pytestmark = pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built")
L = RB.L


def _cloud(n=800, ss=(24, 24, 16), batch=2, seed=0):
    rng = np.random.RandomState(seed)
    c = np.concatenate([np.concatenate([(rng.rand(n, 3) * np.array(ss) * 0.6).astype(np.int64),
                                        np.full((n, 1), b)], 1) for b in range(batch)])
    return c


def test_conv_ops_match_reference():
    S = RB.scn_ref()
    ss = [24, 24, 16]
    coords = _cloud()
    md = S.Metadata_3()
    torch.manual_seed(0)
    feats = torch.randn(len(coords), 5)
    x = feats.new()
    S.InputLayer_updateOutput(md, L(ss), torch.from_numpy(coords), feats, x, 0, 4)
    loc0, prow, header, table = O.input_layer_rules(coords, 4)
    assert np.array_equal(md.getSpatialLocations(L(ss)).numpy(), loc0)
    assert torch.allclose(O.input_layer_forward(feats, header, table), x, atol=1e-6)
    # submanifold 3^3
    w = torch.randn(27, 1, 5, 7) * 0.2
    y = x.new()
    S.SubmanifoldConvolution_updateOutput(L(ss), L([3, 3, 3]), md, x, y, w, torch.Tensor())
    rules = O.submanifold_rules(loc0, ss, [3, 3, 3])
    assert torch.allclose(O.conv_forward(x, w, rules, len(loc0)), y, atol=1e-5)
    dy = torch.randn_like(y)
    dx, dw = x.new(), torch.zeros_like(w)
    S.SubmanifoldConvolution_backward(L(ss), L([3, 3, 3]), md, x, dx, dy, w, dw, torch.Tensor())
    odx, odw, _ = O.conv_backward(x, dy, w, rules)
    assert torch.allclose(odx, dx, atol=1e-5) and torch.allclose(odw, dw, atol=1e-4)
    # strided 2/2 and its deconvolution: compare in canonical row order
    ss1 = [12, 12, 8]
    w2 = torch.randn(8, 1, 5, 6) * 0.3
    y2 = x.new()
    S.Convolution_updateOutput(L(ss), L(ss1), L([2, 2, 2]), L([2, 2, 2]), md, x, y2, w2, torch.Tensor())
    ref_loc1 = md.getSpatialLocations(L(ss1)).numpy()
    loc1, rules2 = O.conv_rules(loc0, ss, [2, 2, 2], [2, 2, 2], ss1)
    po, pr = np.argsort(O.canonical_rank(loc1, ss1)), np.argsort(O.canonical_rank(ref_loc1, ss1))
    assert np.array_equal(loc1[po], ref_loc1[pr])
    oy2 = O.conv_forward(x, w2, rules2, len(loc1))
    assert torch.allclose(oy2[po], y2[pr], atol=1e-5)
    w3 = torch.randn(8, 1, 6, 4) * 0.3
    z = x.new()
    S.Deconvolution_updateOutput(L(ss1), L(ss), L([2, 2, 2]), L([2, 2, 2]), md, y2, z, w3, torch.Tensor())
    oz = O.conv_forward(oy2, w3, rules2, len(loc0), swap=True)
    assert torch.allclose(oz, z, atol=1e-5)
    dz = torch.randn_like(z)
    dy2, dw3 = x.new(), torch.zeros_like(w3)
    S.Deconvolution_backward(L(ss1), L(ss), L([2, 2, 2]), L([2, 2, 2]), md, y2, dy2, dz, w3, dw3, torch.Tensor())
    ody2, odw3, _ = O.conv_backward(oy2, dz, w3, rules2, swap=True)
    assert torch.allclose(ody2[po], dy2[pr], atol=1e-5) and torch.allclose(odw3, dw3, atol=1e-4)
    # SparseToDense
    dense = x.new()
    S.SparseToDense_updateOutput(L(ss1), md, y2, dense, 6)
    assert torch.allclose(O.sparse_to_dense(oy2, loc1, ss1, 2), dense, atol=1e-6)


@pytest.mark.parametrize("leak", [0.0, 0.333, 1.0])
def test_bn_matches_reference(leak):
    S = RB.scn_ref()
    torch.manual_seed(1)
    x = torch.randn(500, 12) * 2 + 0.5
    w, b = torch.rand(12) + 0.5, torch.randn(12)
    rm, rv = torch.zeros(12), torch.ones(12)
    rm2, rv2 = rm.clone(), rv.clone()
    y, sm, si = x.new(), x.new(12), x.new(12)
    S.BatchNormalization_updateOutput(x, y, sm, si, rm, rv, w, b, 1e-4, 0.95, True, leak)
    oy, osm, osi = O.bn_forward(x, w, b, rm2, rv2, 1e-4, 0.95, True, leak)
    assert torch.allclose(oy, y, atol=1e-5) and torch.allclose(osm, sm, atol=1e-6)
    assert torch.allclose(osi, si, rtol=1e-5) and torch.allclose(rm2, rm, atol=1e-6)
    assert torch.allclose(rv2, rv, rtol=1e-5)
    dy = torch.randn_like(y)
    dx, dw, db = x.new(), torch.zeros(12), torch.zeros(12)
    S.BatchNormalization_backward(x, dx, y, dy.clone(), sm, si, rm, rv, w, b, dw, db, leak)
    odx, odw, odb = O.bn_backward(x, oy, dy, osm, osi, w, leak)
    assert torch.allclose(odx, dx, atol=1e-4) and torch.allclose(odw, dw, atol=1e-3)
    assert torch.allclose(odb, db, atol=1e-4)
    # eval mode with given statistics
    y2 = x.new()
    S.BatchNormalization_updateOutput(x, y2, sm, si, rm, rv, w, b, 1e-4, 0.95, False, leak)
    oy2, _, _ = O.bn_forward(x, w, b, rm2, rv2, 1e-4, 0.95, False, leak)
    assert torch.allclose(oy2, y2, atol=1e-5)


def _small_cfg():
    return dict(full_scale=[512, 512, 512], n_planes=[8, 16, 16, 16, 16, 16, 16, 16, 16],
                rpn_map_sizes=[[32, 32, 32], [16, 16, 16], [8, 8, 8], [4, 4, 4]])


def test_ref_backbone_driver_matches_real_fpn_net(gold):
    """oracle/ref_backbone.py (what travels to the GPU box) == the reference's own fpn_net.py"""
    g = gold("small_net")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    net = RB.RefBackbone(sd, **_small_cfg())
    rpn, roi = net.forward(torch.from_numpy(g["locs"]), torch.from_numpy(g["feats"]))
    loss = RB.backbone_loss(rpn, roi)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for i, m in enumerate(list(rpn) + list(roi)):
        loc = m.locations().numpy()
        order = np.argsort(O.canonical_rank(loc, m.ss.tolist()))
        assert np.array_equal(loc[order], g["out%d_loc" % i])
        np.testing.assert_allclose(m.features.detach().numpy()[order], g["out%d_feat" % i], rtol=1e-5, atol=1e-6)
    grads = net.grads()
    n = 0
    for k in g.files:
        if k.startswith("grad/"):
            np.testing.assert_allclose(grads[k[5:]].numpy(), g[k], rtol=1e-4, atol=1e-5)
            n += 1
    assert n > 40
    ge = gold("small_net_eval")
    net.training = False
    with torch.no_grad():
        rpn, roi = net.forward(torch.from_numpy(g["locs"]), torch.from_numpy(g["feats"]))
    for i, m in enumerate(list(rpn) + list(roi)):
        order = np.argsort(O.canonical_rank(m.locations().numpy(), m.ss.tolist()))
        np.testing.assert_allclose(m.features.numpy()[order], ge["out%d_feat" % i], rtol=1e-5, atol=1e-6)


def _wide_cfg():
    return dict(full_scale=[512, 512, 512], n_planes=[32, 64, 32, 32, 32, 32, 32, 32, 32],
                rpn_map_sizes=[[32, 32, 32], [16, 16, 16], [8, 8, 8], [4, 4, 4]])


def wide_state_dict(g):
    return O.seeded_state_dict({k[6:]: g[k] for k in g.files if k.startswith("shape/")}, int(g["seed"]))


def test_ref_backbone_driver_matches_real_fpn_net_wide(gold):
    """the tensor-core-width fixture (tests/golden/wide_net.npz, generated by the reference's own fpn_net.py
    on seeded parameters): the compiled-reference driver reproduces it, and the float64 OracleBackbone is
    within the fp32 bound of it (pins the restatement that serves as ground truth at full size)"""
    g = gold("wide_net")
    sd = wide_state_dict(g)
    locs, feats = torch.from_numpy(g["locs"].astype(np.int64)), torch.from_numpy(g["feats"])
    net = RB.RefBackbone(sd, **_wide_cfg())
    rpn, roi = net.forward(locs, feats)
    loss = RB.backbone_loss(rpn, roi)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for i, m in enumerate(list(rpn) + list(roi)):
        loc = m.locations().numpy()
        order = np.argsort(O.canonical_rank(loc, m.ss.tolist()))
        assert np.array_equal(loc[order], g["out%d_loc" % i])
        np.testing.assert_allclose(m.features.detach().numpy()[order], g["out%d_feat" % i], rtol=1e-5, atol=1e-6)
    grads = net.grads()
    n = 0
    for k in g.files:
        if k.startswith("grad/"):
            np.testing.assert_allclose(O.subsample(grads[k[5:]].numpy()), g[k], rtol=1e-4, atol=1e-5)
            n += 1
    assert n > 40
    truth = O.OracleBackbone(sd, **_wide_cfg())
    trpn, troi = truth.forward(locs.numpy(), feats)
    for i, (tf, tloc, tss) in enumerate(trpn + troi):
        order = np.argsort(O.canonical_rank(tloc, tss))
        assert np.array_equal(tloc[order], g["out%d_loc" % i])
        ref = g["out%d_feat" % i]
        assert np.abs(tf.detach().numpy()[order] - ref).max() <= 1e-4 * np.abs(ref).max()
