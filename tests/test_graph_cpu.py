"""Host logic of the layer-graph compiler (sparseconvnet/graph.py, fpn_net.py) - no GPU: the flat op list the one-call
executor runs must be the reference's layer graph (56 convolution invocations + 34
BatchNormLeakyReLU per forward, fpn_net.py:60-135 / 168-265), and the dead-branch analysis must find exactly the
branches the reference computes without returning them."""
import collections
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

import sparseconvnet as scn
from sparseconvnet import graph as G

FULL_SCALE = [4096, 4096, 512]
PLANES = [32, 64, 64, 128, 128, 128, 256, 256, 256]
RPN_SIZES = [[256, 256, 32], [128, 128, 16], [64, 64, 8], [32, 32, 4]]


def _net(prune=False, planes=PLANES, full=FULL_SCALE, rpn=RPN_SIZES, m=128):
    net = scn.FPN_Net(full, 3, ["xyz", "color", "normal"], 1, planes, nPlaneM=m, residual_blocks=True,
                      fpn_scales_from_top=[4, 3, 2, 1], roi_scales_from_top=(4, 3),
                      downsample=[[[2, 2, 2]] * 8, [[2, 2, 2]] * 8], rpn_map_sizes=rpn, voxel_scale=50,
                      rpn_3d_2d_selector=[1, 2, 3, 4, 5, 6], bn_momentum=0.95, track_running_stats=False)
    net.prune_dead_branches = prune
    return net


def test_backbone_graph_is_the_reference_layer_list():
    net = _net()
    g = net._layer_graph()
    assert g is not None
    kinds = collections.Counter(o.kind for o in g.ops)
    convs = kinds[G.SUBM] + kinds[G.CONV] + kinds[G.DECONV]
    assert convs == 56 and kinds[G.BNRELU] == 34 and kinds[G.ADD] == 17 and len(g.ops) == 107
    # 1 stem 3^3 9->32, 18 residual 3^3 C->C, 9 lateral 1^3 C->128, 8 merge 3^3 128->128 (submanifold);
    # 8 strided 2^3/2 down + 4 z-collapse [1,1,Z] (convolution); 8 deconvolutions 2^3/2 128->128
    assert kinds[G.SUBM] == 1 + 18 + 9 + 8 and kinds[G.CONV] == 8 + 4 and kinds[G.DECONV] == 8
    by_filter = collections.Counter((o.kind, tuple(o.filter), tuple(o.stride)) for o in g.ops if o.kind in (G.SUBM, G.CONV, G.DECONV))
    assert by_filter[(G.SUBM, (3, 3, 3), (1, 1, 1))] == 27 and by_filter[(G.SUBM, (1, 1, 1), (1, 1, 1))] == 9
    assert by_filter[(G.CONV, (2, 2, 2), (2, 2, 2))] == 8 and by_filter[(G.DECONV, (2, 2, 2), (2, 2, 2))] == 8
    zc = sorted(o.filter[2] for o in g.ops if o.kind == G.CONV and tuple(o.filter)[:2] == (1, 1))
    assert zc == [4, 8, 16, 32]
    # channel plan of the down path and the 128-wide top-down path
    stem = g.ops[0]
    assert (stem.kind, stem.n_in_planes, stem.n_out_planes) == (G.SUBM, 9, 32)
    down = [(o.n_in_planes, o.n_out_planes) for o in g.ops if o.kind == G.CONV and tuple(o.filter) == (2, 2, 2)]
    assert down == list(zip(PLANES[:-1], PLANES[1:]))
    assert all((o.n_in_planes, o.n_out_planes) == (128, 128) for o in g.ops if o.kind == G.DECONV)
    # 13 spatial sizes: 9 scales + 4 z-collapsed maps
    assert len(g.sizes) == 13
    assert tuple(FULL_SCALE) in g.sizes and (16, 16, 2) in g.sizes and (256, 256, 1) in g.sizes
    # 6 rpn + 2 roi maps returned; ups[3] is both an rpn map and a roi map: 7 distinct values
    assert len(net._graph_rpn) == 6 and len(net._graph_roi) == 2
    assert len(g.outputs) == 7 and set(g.outputs) == set(net._graph_rpn + net._graph_roi)
    # BN save area: {mean, invstd} per BN layer
    assert g.save_floats == 2 * sum(o.n_out_planes for o in g.ops if o.kind == G.BNRELU)
    # every value is produced exactly once, after its inputs
    made = {0}
    for o in g.ops:
        assert o.in0 in made and (o.kind != G.ADD or o.in1 in made)
        assert o.out not in made
        made.add(o.out)


def test_dead_branches_are_the_unreturned_top_down_path():
    net = _net()
    g = net._layer_graph()
    # the reference runs the whole top-down path and returns the maps of fpn_scales_from_top only
    # (fpn_net.py:186-203): the 4 finest up / merge stages and their 4 lateral convolutions feed nothing
    assert g.n_dead_ops == 21
    dead = [o for o in g.ops if o.out not in g.live_values]
    kinds = collections.Counter(o.kind for o in dead)
    # 4 x (BN, deconvolution, lateral 1^3, add, merge 3^3) at the four finest scales + the z-collapse of the coarsest
    # rpn map, which rpn_3d_2d_selector = [1..6] does not pick
    assert kinds[G.DECONV] == 4 and kinds[G.BNRELU] == 4 and kinds[G.ADD] == 4 and kinds[G.SUBM] == 4 + 4
    assert kinds[G.CONV] == 1 and [tuple(o.filter) for o in dead if o.kind == G.CONV] == [(1, 1, 4)]
    dead_sizes = set(tuple(o.out_ss) for o in dead if o.kind != G.CONV)
    assert dead_sizes == {(4096, 4096, 512), (2048, 2048, 256), (1024, 1024, 128), (512, 512, 64), (256, 256, 32)}
    # (61 % of the forward multiply-adds at the benchmark building's pair counts, SURVEY 8a row G: the dead convolutions
    # are the 128-wide ones of the finest scales)
    assert all(o.n_out_planes == 128 for o in dead)
    pruned = _net(prune=True)._layer_graph()
    assert len(pruned.ops) == 107 - 21 and pruned.outputs == g.outputs
    assert [tuple(o.out_ss) for o in pruned.ops] == [tuple(o.out_ss) for o in g.ops if o.out in g.live_values]
    # parameters of dead layers get no gradient writer in the pruned graph, but stay in the module
    assert len(list(net.parameters())) == len(list(_net(prune=True).parameters()))


def test_gradient_writers_follow_the_reverse_sweep():
    net = _net()
    g = net._layer_graph()
    last = g.param_write_op()
    params = {id(p): n for n, p in net.named_parameters()}
    # every parameter the graph touches has exactly one writer op; the stem is written last (lowest op index)
    assert set(last) <= set(params)
    assert min(last.values()) == 0 and params[[k for k, v in last.items() if v == 0][0]] == "layers_in.1.weight"
    # parameters the backbone never uses on the path (layers_in_0, layers_out, linear) have none
    unused = [n for n, p in net.named_parameters() if id(p) not in last]
    assert unused and all(n.split(".")[0] in ("layers_in_0", "layers_out", "linear") for n in unused), unused


def test_state_dict_keys_equal_the_reference_golden():
    """the key list of the goldens was written by the reference's own fpn_net.py (oracle/make_golden.py)"""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "wide_net.npz"))
    ref_keys = sorted(k[6:] for k in gold.files if k.startswith("shape/"))     # (the wide fixture keeps shapes + a seed)
    assert len(ref_keys) > 150
    net = _net(planes=[32, 64, 32, 32, 32, 32, 32, 32, 32], full=[512] * 3, rpn=[[32] * 3, [16] * 3, [8] * 3, [4] * 3], m=32)
    ours = sorted(net.state_dict().keys())
    assert ours == ref_keys
    for k in ref_keys:
        assert tuple(net.state_dict()[k].shape) == tuple(int(v) for v in gold["shape/" + k]), k


def test_shared_parameters_leave_the_executor():
    """the reverse sweep writes parameter gradients (no accumulation over ops): a weight used by two layers must take
    the per-layer path"""
    conv = scn.SubmanifoldConvolution(3, 8, 8, 3, False)
    g = G.LayerGraph(8, [16, 16, 16])
    v = g.emit(conv, 0)
    v = g.emit(conv, v)
    with pytest.raises(G.Unsupported):
        g.finalize([v])


def test_unsupported_layers_fall_back():
    class Odd(torch.nn.Module):
        pass
    g = G.LayerGraph(8, [16, 16, 16])
    with pytest.raises(G.Unsupported):
        g.emit(Odd(), 0)
    with pytest.raises(G.Unsupported):
        g.emit(scn.SubmanifoldConvolution(2, 8, 8, 3, False), 0)
