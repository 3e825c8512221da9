"""FPN_Net - the sparse3d backbone graph (reference: SparseConvNet/sparseconvnet/fpn_net.py:13-265).

Same constructor signature, same sub-module attribute names (layers_in_0, layers_in, layers_out,
linear, convs_pro2d, m_downs, m_shortcuts, m_ups, m_mergeds => identical state_dict keys) and the
same forward dataflow: 9-scale residual encoder, 1x1x1 shortcuts to nPlaneM, 8 deconvolution
top-down steps with 3^3 merge convs (all 8 are executed, as in the reference), and the
[1,1,Z] z-collapse convolutions that make the 2-D RPN maps.  Debug printing / pdb hooks of the
reference are not reproduced.
"""
import os

import numpy as np
import torch
import torch.nn as nn

import sparseconvnet as scn
from . import graph as _graph

CHECK_NAN = False  # the reference syncs on isnan(weight) every forward (fpn_net.py:141-145)


class FPN_Net(torch.nn.Module):
    def __init__(self, full_scale, dimension, raw_elements, reps, nPlanesF, nPlaneM, residual_blocks,
                 fpn_scales_from_top, roi_scales_from_top, downsample, rpn_map_sizes,
                 rpn_3d_2d_selector, leakiness=0, voxel_scale=None, bn_momentum=0.9,
                 track_running_stats=True):
        nn.Module.__init__(self)
        self.bn_momentum = bn_momentum
        self.track_running_stats = track_running_stats
        self.dimension = dimension
        self.down_kernels, self.down_strides = downsample[0], downsample[1]
        self.fpn_scales_from_top = fpn_scales_from_top
        self.roi_scales_from_top = roi_scales_from_top
        n_scales = len(nPlanesF)
        assert len(self.down_kernels) == n_scales - 1 == len(self.down_strides), \
            "nPlanesF len = %d, kernels num = %d" % (n_scales, len(self.down_kernels))
        assert all(len(k) == 3 for k in self.down_kernels) and all(len(s) == 3 for s in self.down_strides)
        self._merge = "add"
        in_channels = sum({"xyz": 3, "color": 3, "normal": 3}[e] for e in raw_elements)
        bn = dict(momentum=bn_momentum, track_running_stats=track_running_stats)

        self.layers_in_0 = scn.Sequential(scn.InputLayer(dimension, full_scale, mode=4))
        self.layers_in = scn.Sequential(
            scn.InputLayer(dimension, full_scale, mode=4),
            scn.SubmanifoldConvolution(dimension, in_channels, nPlanesF[0], 3, False))
        self.layers_out = scn.Sequential(scn.BatchNormReLU(nPlanesF[0], **bn), scn.OutputLayer(dimension))
        self.linear = nn.Linear(nPlanesF[0], 20)
        self.voxel_scale = voxel_scale
        self.rpn_map_sizes = np.array(rpn_map_sizes)
        self.rpn_3d_2d_selector = rpn_3d_2d_selector

        self.convs_pro2d = nn.ModuleList(
            scn.Convolution(dimension, nPlaneM, nPlaneM, [1, 1, int(z)], [1, 1, 1], False)
            for z in self.rpn_map_sizes[:, -1])

        def bn_act(c):
            return scn.BatchNormLeakyReLU(c, leakiness=leakiness, **bn)

        def block(m, a, b):
            if residual_blocks:
                m.add(scn.ConcatTable()
                      .add(scn.Identity() if a == b else scn.NetworkInNetwork(a, b, False))
                      .add(scn.Sequential()
                           .add(bn_act(a)).add(scn.SubmanifoldConvolution(dimension, a, b, 3, False))
                           .add(bn_act(b)).add(scn.SubmanifoldConvolution(dimension, b, b, 3, False)))
                      ).add(scn.AddTable())
            else:
                m.add(scn.Sequential().add(bn_act(a))
                      .add(scn.SubmanifoldConvolution(dimension, a, b, 3, False)))
            return {"kernel": [1, 1, 1], "stride": [1, 1, 1]}

        def strided(layer, a, b, scale):
            return scn.Sequential().add(bn_act(a)).add(
                layer(dimension, a, b, self.down_kernels[scale], self.down_strides[scale], False))

        self.m_downs, self.m_shortcuts = nn.ModuleList(), nn.ModuleList()
        self.operations_down, self.operations_up = [], []
        for k in range(n_scales):
            m = scn.Sequential()
            if k > 0:
                m.add(strided(scn.Convolution, nPlanesF[k - 1], nPlanesF[k], k - 1))
                self.operations_down.append({"kernel": self.down_kernels[k - 1],
                                             "stride": self.down_strides[k - 1]})
            for _ in range(reps):
                op = block(m, nPlanesF[k], nPlanesF[k])
                if k == 0:
                    self.operations_down.append(op)
            self.m_downs.append(m)
            self.m_shortcuts.append(scn.SubmanifoldConvolution(dimension, nPlanesF[k], nPlaneM, 1, False))

        self.m_ups, self.m_mergeds = nn.ModuleList(), nn.ModuleList()
        for k in range(n_scales - 1, 0, -1):
            up = strided(scn.Deconvolution, nPlaneM, nPlaneM, k - 1)
            # the reference adds BN and Deconvolution directly to the Sequential (fpn_net.py:88-90):
            # keys m_ups.<i>.0 (BN) and m_ups.<i>.1 (Deconvolution) - `strided` builds exactly that
            self.m_ups.append(up)
            self.operations_up.append({"kernel": self.down_kernels[k - 1],
                                       "stride": self.down_strides[k - 1]})
            self.m_mergeds.append(scn.SubmanifoldConvolution(dimension, nPlaneM, nPlaneM, 3, False))

    def forward(self, net0):
        if CHECK_NAN and not torch.isnan(self.layers_in[1].weight).sum() == 0:
            raise FloatingPointError("FPN_Net: NaN in stem weights")
        net = self.layers_in[0](net0)                  # InputLayer: voxel hashing (or a PreparedInput)
        if not getattr(net, "rulebooks_built", False):
            self._prebuild_rulebooks(net)
        g = self._layer_graph() if self.use_layer_graph else None
        mode = g.bn_mode(self.training, self) if g is not None else None
        if mode is not None:
            return self.forward_fpn_graph(net, g, mode)
        return self.forward_fpn(self.layers_in[1](net))

    # ---- one-call execution of the whole graph (sparseconvnet/graph.py) ----------------------
    use_layer_graph = os.environ.get("SCN_B200_LAYER_GRAPH", "1") != "0"
    # B200 extension, OFF by default: skip the layers no returned map depends on (the finer half of the top-down path,
    # which the reference computes and drops - fpn_net.py:186-203).  Outputs and gradients are unchanged bit for bit;
    # with tracked running statistics the BN buffers of the skipped layers are no longer updated.  Set the attribute
    # before the first forward pass (or call `invalidate_graph()` after changing it).
    prune_dead_branches = os.environ.get("SCN_B200_PRUNE_DEAD", "0") == "1"

    def invalidate_graph(self):
        self._graph_cache = None

    def _apply(self, fn, *args, **kwargs):
        self._graph_cache = None          # .to() / .cuda() replace the BN buffers the compiled graph points at
        return super(FPN_Net, self)._apply(fn, *args, **kwargs)

    def _layer_graph(self):
        """the ops of layers_in[1] + forward_fpn as a flat list, compiled once (None if a layer is not
        supported by the executor - the per-layer path below then runs)"""
        g = getattr(self, "_graph_cache", None)
        if g is None:
            try:
                il = self.layers_in[0]
                g = _graph.LayerGraph(self.layers_in[1].nIn, il.spatial_size.tolist())
                v = g.emit(self.layers_in[1], 0)
                n_scales = len(self.m_downs)
                downs = []
                for m in self.m_downs:
                    v = g.emit(m, v)
                    downs.append(v)
                v = g.emit(self.m_shortcuts[-1], v)
                ups = [v]
                for k in range(n_scales - 1):
                    j = n_scales - 2 - k
                    v = g.emit(self.m_ups[k], v)
                    v = g.add(v, g.emit(self.m_shortcuts[j], downs[j]))
                    ups.append(g.emit(self.m_mergeds[k], v))     # the top-down path continues from the SUM
                rpn3d = [ups[i] for i in self.fpn_scales_from_top]
                for i, v3 in enumerate(rpn3d):
                    assert list(g.values[v3][1]) == [int(x) for x in self.rpn_map_sizes[i]]
                rpn2d = [g.emit(self.convs_pro2d[i], rpn3d[i]) for i in range(len(rpn3d))]
                both = rpn3d + rpn2d
                self._graph_rpn = [both[i] for i in self.rpn_3d_2d_selector]
                self._graph_roi = [ups[i] for i in self.roi_scales_from_top]
                g.finalize(self._graph_rpn + self._graph_roi, prune_dead=self.prune_dead_branches)
            except _graph.Unsupported:
                g = False
            self._graph_cache = g
        if g:
            g.grad_sink = getattr(self, "_grad_sink", None)     # a GradBucket built with module=self
        return g or None

    def forward_fpn_graph(self, net, g, bn_mode):
        feats = _graph.GraphFunction.apply(g, net.metadata, bn_mode, net.features, *g.grad_params)
        by_value = dict(zip(g.outputs, feats))

        def wrap(v):
            t = scn.SparseConvNetTensor(metadata=net.metadata, spatial_size=torch.tensor(g.values[v][1]))
            t.features = by_value[v]
            return t
        return [wrap(v) for v in self._graph_rpn], [wrap(v) for v in self._graph_roi]

    def prepare(self, coords, batch_size=0):
        """All integer work of a batch - input grid, the 12 coarser grids and every rulebook - on the
        current stream; returns a PreparedInput to pass as `net([prepared, features])`.  Meant to run
        one batch ahead on a side stream (scn.InputPrefetcher(net.prepare))."""
        il = self.layers_in[0]
        m = scn.Metadata(self.dimension)
        n_active = scn.SCN.InputLayer_prepare_plan(m, il.spatial_size, coords, batch_size, il.mode,
                                                   self._rulebook_plan(il.spatial_size))
        ev = torch.cuda.Event()
        ev.record()
        p = scn.PreparedInput(m, il.spatial_size, n_active, coords.size(0), ev, coords)
        p.rulebooks_built = True
        return p

    def _rulebook_plan(self, ss):
        """[n_ops, 13] int64: (kind, in_size, out_size, filter, stride) of every rulebook of the graph"""
        key = tuple(ss.tolist())
        if getattr(self, "_plan_key", None) != key:
            ops, three, ones = [], [3] * self.dimension, [1] * self.dimension
            cur = list(key)
            for k in range(len(self.m_downs)):
                ops.append([0] + cur + cur + three + ones)
                if k + 1 < len(self.m_downs):
                    fs, st = list(self.down_kernels[k]), list(self.down_strides[k])
                    out = [(c - f) // s + 1 for c, f, s in zip(cur, fs, st)]
                    ops.append([1] + cur + out + fs + st)
                    cur = out
            for conv, size in zip(self.convs_pro2d, self.rpn_map_sizes):
                size = [int(v) for v in size]
                fs, st = conv.filter_size.tolist(), conv.filter_stride.tolist()
                ops.append([1] + size + [(c - f) // s + 1 for c, f, s in zip(size, fs, st)] + fs + st)
            self._plan, self._plan_key = torch.tensor(ops, dtype=torch.int64), key
        return self._plan

    def _prebuild_rulebooks(self, net):
        """Build every hash grid / rulebook the graph below will ask for, back to back, before any
        feature kernel is queued.  Each build needs a count read-back; done lazily (as the reference
        does, layer by layer) those read-backs would drain the GPU between layers."""
        m, ss = net.metadata, net.spatial_size
        three = torch.tensor([3] * self.dimension)
        for k in range(len(self.m_downs)):
            m.prepareSubmanifoldRuleBook(ss, three)
            if k + 1 < len(self.m_downs):
                fs, st = torch.tensor(self.down_kernels[k]), torch.tensor(self.down_strides[k])
                out = (ss - fs) // st + 1
                m.prepareRuleBook(ss, out, fs, st)
                ss = out
        for conv, size in zip(self.convs_pro2d, self.rpn_map_sizes):
            size = torch.tensor([int(v) for v in size])
            m.prepareRuleBook(size, (size - conv.filter_size) // conv.filter_stride + 1, conv.filter_size,
                              conv.filter_stride)

    def forward_fpn(self, net):
        n_scales = len(self.m_downs)
        downs = []
        for m in self.m_downs:
            net = m(net)
            downs.append(net)
        net = self.m_shortcuts[-1](net)
        ups = [net]
        for k in range(n_scales - 1):
            j = n_scales - 2 - k
            net = self.m_ups[k](net)
            net = scn.add_feature_planes([net, self.m_shortcuts[j](downs[j])])
            ups.append(self.m_mergeds[k](net))

        rpn_maps_3d = [ups[i] for i in self.fpn_scales_from_top]
        rpn_maps_2d = [self.convs_pro2d[i](rpn_maps_3d[i]) for i in range(len(rpn_maps_3d))]
        both = rpn_maps_3d + rpn_maps_2d
        rpn_maps = [both[i] for i in self.rpn_3d_2d_selector]
        roi_maps = [ups[i] for i in self.roi_scales_from_top]
        for i, t in enumerate(rpn_maps_3d):
            assert torch.all(t.spatial_size == torch.tensor(self.rpn_map_sizes[i]))
        return rpn_maps, roi_maps
