"""The oracle (oracle/scn_oracle.py) against the compiled reference (oracle/_ref/SCN_ref.so = the reference's own
pybind.cpp + sparseconvnet_cpu.cpp, unmodified) over a sweep of clouds and layer geometries: ragged batches with an
empty sample, single points, heavy duplication, every InputLayer mode, overlapping strides, z-collapse filters,
NetworkInNetwork, OutputLayer and SparseToDense gradients.  Integer results bit-exact (after the canonical relabelling
for conv-created scales, whose numbering the reference leaves to hash iteration order), features to fp32 rounding."""
import numpy as np
import pytest
import torch

import ref_backbone as RB
import scn_oracle as O

pytestmark = pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built")
L = RB.L
E = torch.Tensor


def _cloud(sizes, ss, seed, dup=0.0):
    """one block of points per sample (sizes[b] points; 0 = an empty sample in the middle of the batch)"""
    rng = np.random.RandomState(seed)
    parts = []
    for b, n in enumerate(sizes):
        c = (rng.rand(n, 3) * np.array(ss) * 0.7).astype(np.int64)
        if n >= 6:
            c[: n // 3, 2] = ss[2] // 3                              # a populated plane
        if dup and n:
            c = np.concatenate([c, c[rng.randint(0, n, int(n * dup))]])
        parts.append(np.concatenate([c, np.full((len(c), 1), b)], 1))
    return np.concatenate(parts).astype(np.int64)


CLOUDS = [
    ([400, 300], (24, 24, 16), 0.0),
    ([150, 0, 220, 1], (16, 20, 12), 0.3),          # empty sample, single-point sample, duplicates
    ([1], (8, 8, 8), 0.0),
    ([700], (12, 12, 6), 2.0),                      # every voxel hit several times
]


@pytest.mark.parametrize("ci", range(len(CLOUDS)))
@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_input_layer_modes_match_reference(ci, mode):
    sizes, ss, dup = CLOUDS[ci]
    S = RB.scn_ref()
    coords = _cloud(sizes, ss, 10 + ci, dup)
    torch.manual_seed(ci)
    feats = torch.randn(len(coords), 4)
    md = S.Metadata_3()
    x = feats.new()
    S.InputLayer_updateOutput(md, L(ss), torch.from_numpy(coords), feats, x, len(sizes), mode)
    loc0, prow, header, table = O.input_layer_rules(coords, mode)
    assert np.array_equal(md.getSpatialLocations(L(ss)).numpy(), loc0)
    assert torch.allclose(O.input_layer_forward(feats, header, table), x, atol=1e-6)
    # OutputLayer: every point gets its voxel's row back (CPU/IOLayers.cpp:73-104)
    out = feats.new()
    S.OutputLayer_updateOutput(md, x, out)
    if mode in (3, 4):
        assert torch.equal(out, x[torch.from_numpy(prow)])


@pytest.mark.parametrize("ci", range(len(CLOUDS)))
@pytest.mark.parametrize("fs", [[3, 3, 3], [1, 1, 1], [3, 1, 3], [5, 5, 5], [2, 2, 2]])
def test_submanifold_rules_and_features_match_reference(ci, fs):
    sizes, ss, dup = CLOUDS[ci]
    S, D = RB.scn_ref(), RB.scn_refdump()
    coords = _cloud(sizes, ss, 20 + ci, dup)
    torch.manual_seed(ci)
    md = S.Metadata_3()
    x = torch.Tensor()
    S.InputLayer_updateOutput(md, L(ss), torch.from_numpy(coords), torch.randn(len(coords), 3), x, len(sizes), 4)
    loc0 = O.input_layer_rules(coords, 4)[0]
    K = fs[0] * fs[1] * fs[2]
    w = torch.randn(K, 1, 3, 5) * 0.3
    y = x.new()
    S.SubmanifoldConvolution_updateOutput(L(ss), L(fs), md, x, y, w, E())
    rules = O.submanifold_rules(loc0, ss, fs)
    assert torch.allclose(O.conv_forward(x, w, rules, len(loc0)), y, atol=1e-5)
    dy = torch.randn_like(y)
    dx, dw = x.new(), torch.zeros_like(w)
    S.SubmanifoldConvolution_backward(L(ss), L(fs), md, x, dx, dy, w, dw, E())
    odx, odw, _ = O.conv_backward(x, dy, w, rules)
    assert torch.allclose(odx, dx, atol=1e-5) and torch.allclose(odw, dw, atol=1e-4)


GEOM = [([2, 2, 2], [2, 2, 2]), ([3, 3, 3], [2, 2, 2]), ([1, 1, 0], [1, 1, 1]), ([2, 2, 1], [2, 2, 1]), ([4, 4, 4], [2, 2, 2])]


@pytest.mark.parametrize("ci", range(len(CLOUDS)))
@pytest.mark.parametrize("gi", range(len(GEOM)))
def test_strided_rules_and_features_match_reference(ci, gi):
    sizes, ss, dup = CLOUDS[ci]
    fs, st = [list(v) for v in GEOM[gi]]
    if fs[2] == 0:
        fs[2] = ss[2]                                                # z-collapse [1,1,Z]
    out_ss = [(s - f) // t + 1 for s, f, t in zip(ss, fs, st)]
    assert min(out_ss) >= 1
    # the input extent a filter / stride pair tiles exactly (convolution.py:37-38 asserts it): at most stride - 1 less
    # than the cloud's, and the points only fill 70 % of it
    ss = [(o - 1) * t + f for o, t, f in zip(out_ss, st, fs)]
    S = RB.scn_ref()
    coords = _cloud(sizes, ss, 30 + ci, dup)
    torch.manual_seed(ci + gi)
    md = S.Metadata_3()
    x = torch.Tensor()
    S.InputLayer_updateOutput(md, L(ss), torch.from_numpy(coords), torch.randn(len(coords), 4), x, len(sizes), 4)
    loc0 = O.input_layer_rules(coords, 4)[0]
    K = fs[0] * fs[1] * fs[2]
    w = torch.randn(K, 1, 4, 6) * 0.3
    y = x.new()
    S.Convolution_updateOutput(L(ss), L(out_ss), L(fs), L(st), md, x, y, w, E())
    ref_loc1 = md.getSpatialLocations(L(out_ss)).numpy()
    loc1, rules = O.conv_rules(loc0, ss, fs, st, out_ss)
    po, pr = np.argsort(O.canonical_rank(loc1, out_ss)), np.argsort(O.canonical_rank(ref_loc1, out_ss))
    assert np.array_equal(loc1[po], ref_loc1[pr])                    # the same sites
    assert (np.diff(ref_loc1[:, 3]) >= 0).all() and (np.diff(loc1[:, 3]) >= 0).all()     # batch-contiguous, both
    oy = O.conv_forward(x, w, rules, len(loc1))
    assert torch.allclose(oy[po], y[pr], atol=1e-5)
    dy = torch.randn_like(y)
    ody = torch.empty_like(dy)
    ody[po] = dy[pr]                                                 # the same gradient, in the oracle's row order
    dx, dw = x.new(), torch.zeros_like(w)
    S.Convolution_backward(L(ss), L(out_ss), L(fs), L(st), md, x, dx, dy, w, dw, E())
    odx, odw, _ = O.conv_backward(x, ody, w, rules)
    assert torch.allclose(odx, dx, atol=1e-5) and torch.allclose(odw, dw, atol=1e-4)
    # deconvolution back to the fine scale: the same rulebook with the roles swapped (CPU/Deconvolution.cpp:15-16)
    w2 = torch.randn(K, 1, 6, 3) * 0.3
    z = x.new()
    S.Deconvolution_updateOutput(L(out_ss), L(ss), L(fs), L(st), md, y, z, w2, E())
    assert torch.allclose(O.conv_forward(oy, w2, rules, len(loc0), swap=True), z, atol=1e-5)
    # SparseToDense of the coarse map and its gradient (CPU/SparseToDense.cpp:8-87)
    dense = x.new()
    S.SparseToDense_updateOutput(L(out_ss), md, y, dense, 6)
    assert torch.allclose(O.sparse_to_dense(oy, loc1, out_ss, len(sizes)), dense, atol=1e-6)
    dd = torch.randn_like(dense)
    dsp = x.new()
    S.SparseToDense_updateGradInput(L(out_ss), md, y, dsp, dd)
    c = torch.from_numpy(ref_loc1)
    assert torch.equal(dsp, dd[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]])


def test_network_in_network_matches_reference():
    S = RB.scn_ref()
    torch.manual_seed(5)
    x, w, b = torch.randn(300, 6), torch.randn(6, 9), torch.randn(9)
    y = x.new()
    S.NetworkInNetwork_updateOutput(x, y, w, b)
    assert torch.allclose(y, x @ w + b, atol=1e-5)
    dy = torch.randn_like(y)
    dx = x.new()
    S.NetworkInNetwork_updateGradInput(dx, dy, w)
    assert torch.allclose(dx, dy @ w.t(), atol=1e-5)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    S.NetworkInNetwork_accGradParameters(x, dy, dw, db)
    assert torch.allclose(dw, x.t() @ dy, atol=1e-4) and torch.allclose(db, dy.sum(0), atol=1e-4)
    # the layer graph runs it as a 1^3 submanifold convolution: the same numbers through the oracle's rules
    loc = np.concatenate([np.arange(300)[:, None] % 7, np.arange(300)[:, None] // 7 % 7, np.arange(300)[:, None] // 49,
                          np.zeros((300, 1), np.int64)], 1)
    rules = O.submanifold_rules(loc, [7, 7, 7], [1, 1, 1])
    assert torch.allclose(O.conv_forward(x, w.view(1, 1, 6, 9), rules, 300, bias=b), y, atol=1e-5)


@pytest.mark.parametrize("n", [2, 3, 1000])
def test_bn_small_row_counts_match_reference(n):
    """n = 2 / 3: the unbiased running-variance divides by n - 1 (CPU/BatchNormalization.cpp:37-38)"""
    S = RB.scn_ref()
    torch.manual_seed(n)
    C = 8
    x = torch.randn(n, C)
    w, b = torch.rand(C) + 0.5, torch.randn(C)
    rm, rv = torch.randn(C), torch.rand(C) + 0.5
    rm2, rv2 = rm.clone(), rv.clone()
    y, sm, si = x.new(), x.new(C), x.new(C)
    S.BatchNormalization_updateOutput(x, y, sm, si, rm, rv, w, b, 1e-4, 0.9, True, 0.0)
    oy, osm, osi = O.bn_forward(x, w, b, rm2, rv2, 1e-4, 0.9, True, 0.0)
    assert torch.allclose(oy, y, atol=1e-5) and torch.allclose(osm, sm, atol=1e-6) and torch.allclose(osi, si, rtol=1e-4)
    assert torch.allclose(rm2, rm, atol=1e-6) and torch.allclose(rv2, rv, rtol=1e-4, atol=1e-6)
