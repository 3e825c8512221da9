"""Layer-graph execution of scn module trees (B200 extension; C side: csrc/graph.cu, include/scn_b200.h
"layer-graph executor").

The reference runs a network as one Python autograd Function per layer (submanifoldConvolution.py:60-113,
convolution.py:64-126, deconvolution.py:64-155, batchNormalization.py:70-187, composed by sequential.py:15-17
and tables.py:28-56).  At batch 1 that is ~260 Function calls per step and the main thread's enqueue work,
not the GPU, bounds the step.  `LayerGraph` compiles a tree of scn modules ONCE into a flat op list;
`GraphFunction` then runs the whole list forward in one foreign call and backward in another, through
exactly the per-op C entry points the layer files call (same kernels, bit-identical results).  Parameters
stay ordinary torch Parameters: they are inputs of the one autograd Function, so optimizers, DDP hooks and
state_dicts see nothing different.
"""
import ctypes
from ctypes import POINTER, c_double, c_float, c_int32, c_int64, c_uint8, c_void_p

import torch
from torch.autograd import Function

import sparseconvnet
from . import _lib
from . import modules as M


class Unsupported(Exception):
    """the module tree contains a layer the executor does not take (callers fall back to per-layer calls)"""


class GraphOp(ctypes.Structure):   # include/scn_b200.h: scn_graph_op_t
    _fields_ = [("kind", c_int32), ("in0", c_int32), ("in1", c_int32), ("out", c_int32),
                ("n_in_planes", c_int32), ("n_out_planes", c_int32),
                ("p0", c_int32), ("p1", c_int32), ("p2", c_int32), ("p3", c_int32),
                ("in_ss", c_int64 * 3), ("out_ss", c_int64 * 3), ("filter", c_int64 * 3), ("stride", c_int64 * 3),
                ("save_off", c_int64),
                ("eps", c_float), ("momentum", c_float), ("leakiness", c_float), ("pad_", c_float)]


SUBM, CONV, DECONV, BNRELU, ADD = 1, 2, 3, 4, 5


class LayerGraph(object):
    def __init__(self, in_planes, spatial_size):
        self.values = [(int(in_planes), tuple(int(v) for v in spatial_size))]   # value 0 = the input features
        self.ops = []
        self.params = []          # tensors: Parameters and BN running statistics, by slot
        self._slot = {}
        self.save_floats = 0
        self.outputs = []
        self._c_ops = None
        self.grad_sink = None     # GradBucket.attach(): parameter gradients are written straight into the bucket

    def param_write_op(self):
        """{id(Parameter): index of the LAST op (lowest index = latest in the reverse sweep) that writes its
        gradient}: what a gradient bucket needs to know when a range of it is final"""
        last = {}
        for i, o in enumerate(self.ops):
            for slot in ((o.p0, o.p1) if o.kind != ADD else ()):
                if slot >= 0 and isinstance(self.params[slot], torch.nn.Parameter):
                    k = id(self.params[slot])
                    last[k] = min(last.get(k, i), i)
        return last

    # ---- construction ---------------------------------------------------------------------
    def _param(self, t):
        if t is None:
            return -1
        s = self._slot.get(id(t))
        if s is None:
            s = self._slot[id(t)] = len(self.params)
            self.params.append(t)
        return s

    def _value(self, planes, ss):
        self.values.append((int(planes), tuple(int(v) for v in ss)))
        return len(self.values) - 1

    def _op(self, kind, in0, out, cin, cout, in_ss, out_ss, filt=(1, 1, 1), stride=(1, 1, 1), p=(-1, -1, -1, -1),
            in1=-1, eps=0.0, momentum=0.0, leak=0.0, save_off=0):
        o = GraphOp()
        o.kind, o.in0, o.in1, o.out = kind, in0, in1, out
        o.n_in_planes, o.n_out_planes = cin, cout
        o.p0, o.p1, o.p2, o.p3 = p
        for i in range(3):
            o.in_ss[i], o.out_ss[i], o.filter[i], o.stride[i] = int(in_ss[i]), int(out_ss[i]), int(filt[i]), int(stride[i])
        o.save_off, o.eps, o.momentum, o.leakiness = save_off, eps, momentum, leak
        self.ops.append(o)
        return out

    def add(self, a, b):
        (ca, sa), (cb, sb) = self.values[a], self.values[b]
        assert ca == cb and sa == sb, "add of values with different shapes"
        return self._op(ADD, a, self._value(ca, sa), ca, ca, sa, sa, in1=b)

    def emit(self, mod, v):
        """append the ops of module `mod` applied to value `v`; returns the output value"""
        planes, ss = self.values[v]
        if isinstance(mod, M.Identity):
            return v
        if isinstance(mod, M.BatchNormalization):
            assert mod.nPlanes == planes, (mod, planes)
            p = (self._param(getattr(mod, "weight", None)), self._param(getattr(mod, "bias", None)),
                 self._param(mod.running_mean), self._param(mod.running_var))
            off = self.save_floats
            self.save_floats += 2 * planes
            return self._op(BNRELU, v, self._value(planes, ss), planes, planes, ss, ss, p=p, eps=mod.eps,
                            momentum=mod.momentum, leak=mod.leakiness, save_off=off)
        if isinstance(mod, (M.SubmanifoldConvolution, M.Convolution, M.Deconvolution)):
            if mod.groups != 1 or mod.dimension != 3:
                raise Unsupported(repr(mod))
            assert mod.nIn == planes, (mod, planes)
            p = (self._param(mod.weight), self._param(getattr(mod, "bias", None)), -1, -1)
            fs = mod.filter_size.tolist()
            if isinstance(mod, M.SubmanifoldConvolution):
                return self._op(SUBM, v, self._value(mod.nOut, ss), planes, mod.nOut, ss, ss, fs, p=p)
            st = mod.filter_stride.tolist()
            if isinstance(mod, M.Convolution):
                out_ss = [(c - f) // s + 1 for c, f, s in zip(ss, fs, st)]
                assert all((o - 1) * s + f == c for o, s, f, c in zip(out_ss, st, fs, ss)), (ss, fs, st)
                return self._op(CONV, v, self._value(mod.nOut, out_ss), planes, mod.nOut, ss, out_ss, fs, st, p=p)
            out_ss = [(c - 1) * s + f for c, f, s in zip(ss, fs, st)]
            return self._op(DECONV, v, self._value(mod.nOut, out_ss), planes, mod.nOut, ss, out_ss, fs, st, p=p)
        if isinstance(mod, M.NetworkInNetwork):
            # a 1x1x1 submanifold convolution: identical memory layout of weight / gradient
            p = (self._param(mod.weight), self._param(getattr(mod, "bias", None)), -1, -1)
            return self._op(SUBM, v, self._value(mod.nOut, ss), planes, mod.nOut, ss, ss, (1, 1, 1), p=p)
        if isinstance(mod, (M.Sequential, torch.nn.Sequential)) and not isinstance(mod, M._Table):
            cur = v
            for child in mod._modules.values():
                if isinstance(child, M.ConcatTable):
                    if isinstance(cur, list):
                        raise Unsupported("nested tables")
                    cur = [self.emit(c, cur) for c in child._modules.values()]
                elif isinstance(child, M.AddTable):
                    if not isinstance(cur, list):
                        raise Unsupported("AddTable without ConcatTable")
                    acc = cur[0]
                    for other in cur[1:]:
                        acc = self.add(acc, other)
                    cur = acc
                else:
                    if isinstance(cur, list):
                        raise Unsupported("ConcatTable not followed by AddTable")
                    cur = self.emit(child, cur)
            if isinstance(cur, list):
                raise Unsupported("dangling ConcatTable")
            return cur
        raise Unsupported(type(mod).__name__)

    def finalize(self, outputs, prune_dead=False):
        """prune_dead: drop the ops no output depends on.  The reference computes them (fpn_net.py:186-203 runs the whole
        top-down path and returns only the maps of `fpn_scales_from_top`: 61 % of the forward MACs at the BASELINE
        config) and so does this library by default; nothing observable depends on them except the running statistics
        of their BN layers when those are tracked - outputs and gradients are bit-identical with and without."""
        self.outputs = list(dict.fromkeys(outputs))          # unique, order kept
        assert 0 not in self.outputs
        live = set(self.outputs)
        for o in reversed(self.ops):
            if o.out in live:
                live.add(o.in0)
                if o.kind == ADD:
                    live.add(o.in1)
        self.n_dead_ops = sum(1 for o in self.ops if o.out not in live)
        self.live_values = live
        if prune_dead:
            self.ops = [o for o in self.ops if o.out in live]
        self._c_ops = (GraphOp * len(self.ops))(*self.ops)
        self.grad_params = [t for t in self.params if isinstance(t, torch.nn.Parameter)]
        self._grad_slot = [self._slot[id(t)] for t in self.grad_params]
        self.sizes = sorted(set(ss for _, ss in self.values))
        self.is_weight = [False] * len(self.params)
        writers = {}
        for o in self.ops:
            if o.kind in (SUBM, CONV, DECONV):
                self.is_weight[o.p0] = True
            if o.kind != ADD:
                for slot in (o.p0, o.p1):
                    if slot >= 0 and isinstance(self.params[slot], torch.nn.Parameter):
                        writers[slot] = writers.get(slot, 0) + 1
        if any(c > 1 for c in writers.values()):
            # the reverse sweep WRITES parameter gradients (no accumulation over ops): a parameter shared by two
            # layers runs through the per-layer Functions instead
            raise Unsupported("a parameter is shared by several layers")
        return self

    def bn_mode(self, training, root):
        """BN mode of one graph execution (scn_batchnorm_forward `train`): 1 training, 0 evaluation with the
        running buffers, 2 evaluation with batch statistics (track_running_stats=False,
        batchNormalization.py:51-56); None when the BN layers of `root` disagree (per-layer path then)"""
        if training:
            return 1
        track = set(bool(m.track_running_stats) for m in root.modules() if isinstance(m, M.BatchNormalization))
        if len(track) > 1:
            return None
        return 0 if (not track or track.pop()) else 2

    # ---- execution --------------------------------------------------------------------------
    def _param_arrays(self, plist):
        """device pointers and weight tags by slot; `plist` = the tensors autograd handed to forward for
        grad_params (same storage as the Parameters)"""
        n = len(self.params)
        ptrs = (c_void_p * n)()
        tags = (c_int64 * (2 * n))()
        for i, t in enumerate(self.params):
            ptrs[i] = t.data_ptr()
            if self.is_weight[i]:
                tag = _lib.weight_tag(t)
                if tag is not None:
                    tags[2 * i], tags[2 * i + 1] = tag[0], tag[1]
        return ptrs, tags


class GraphFunction(Function):
    """forward(graph, metadata, train, x0, *graph.grad_params) -> one feature tensor per graph output"""

    @staticmethod
    def forward(ctx, graph, metadata, train, x0, *plist):
        x0 = _lib.require_cuda_f32(x0, "input_features").contiguous()
        for t in graph.params:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("sparseconvnet (B200): graph parameters must be contiguous CUDA float32 tensors")
        n_act = {ss: metadata.getNActive(list(ss)) for ss in graph.sizes}
        for ss, n in n_act.items():
            if n < 0:
                raise RuntimeError("Metadata: no grid at spatial size %s" % (list(ss),))
        nv = len(graph.values)
        rows = (c_int64 * nv)(*[n_act[ss] for _, ss in graph.values])
        if rows[0] != x0.size(0):
            raise RuntimeError("input features have %d rows, the input grid %d" % (x0.size(0), rows[0]))
        is_out = set(graph.outputs)
        off, total = [0] * nv, 0
        written = set(o.out for o in graph.ops)
        for v in range(1, nv):
            if v not in is_out and v in written:                              # (pruned graphs: dead values get no room)
                off[v] = total
                total += (rows[v] * graph.values[v][0] + 63) // 64 * 64       # 256-byte aligned
        arena = x0.new_empty(max(total, 1))
        outs = {v: x0.new_empty(rows[v], graph.values[v][0]) for v in graph.outputs}
        base = arena.data_ptr()
        vals = (c_void_p * nv)()
        vals[0] = x0.data_ptr()
        for v in range(1, nv):
            vals[v] = outs[v].data_ptr() if v in is_out else base + 4 * off[v]
        bn_save = x0.new_empty(max(graph.save_floats, 1))
        ptrs, tags = graph._param_arrays(plist)
        macs = c_double()
        _lib.check(_lib.lib.scn_graph_forward(metadata._h, graph._c_ops, len(graph.ops), vals, rows, ptrs, tags,
                                              _lib.ptr(bn_save), int(train), _lib.precision(), _lib.stream(),
                                              ctypes.byref(macs)))
        sparseconvnet.forward_pass_multiplyAdd_count += macs.value
        ctx.graph, ctx.metadata_, ctx.keep = graph, metadata, (x0, arena, bn_save, outs)
        # the backward pass re-reads the live Parameters (not saved copies): remember their version counters so that an
        # in-place edit between forward and backward is an error, as it is for tensors saved by autograd
        ctx.param_versions = [t._version for t in graph.grad_params]
        ctx.rows, ctx.vals, ctx.off, ctx.total = rows, vals, off, total
        ctx.set_materialize_grads(False)
        return tuple(outs[v] for v in graph.outputs)

    @staticmethod
    def backward(ctx, *gouts):
        graph, metadata = ctx.graph, ctx.metadata_
        if ctx.keep is None:
            raise RuntimeError("sparseconvnet layer graph: backward called a second time - the step's value arena is released "
                               "after the first backward (retain_graph=True is not supported; set use_layer_graph = False "
                               "on the network for that)")
        x0, arena, bn_save, outs = ctx.keep
        for t, v in zip(graph.grad_params, ctx.param_versions):
            if t._version != v:
                raise RuntimeError("sparseconvnet layer graph: a parameter was modified in place between the forward and "
                                   "the backward pass (its gradient would be computed with the new values)")
        rows, nv = ctx.rows, len(graph.values)
        # gradient buffers: one arena laid out like the value arena, separate buffers for the outputs
        garena = x0.new_empty(max(ctx.total, 1))
        gbase = garena.data_ptr()
        gout_own = {v: torch.empty_like(outs[v]) for v in graph.outputs}
        grads = (c_void_p * nv)()
        d_x0 = torch.empty_like(x0) if ctx.needs_input_grad[3] else None
        grads[0] = d_x0.data_ptr() if d_x0 is not None else None
        is_out = set(graph.outputs)
        for v in range(1, nv):
            grads[v] = gout_own[v].data_ptr() if v in is_out else gbase + 4 * ctx.off[v]
        ext = (c_void_p * nv)()
        keep = []
        for v, g in zip(graph.outputs, gouts):
            if g is not None:
                g = _lib.require_cuda_f32(g, "grad_output").contiguous()
                keep.append(g)
                ext[v] = g.data_ptr()
        n = len(graph.params)
        numel = [t.numel() for t in graph.params]
        poff, ptotal = [0] * n, 0
        for i in graph._grad_slot:
            poff[i] = ptotal
            ptotal += (numel[i] + 3) // 4 * 4
        # a GradBucket attached to the graph takes the parameter gradients directly (no AccumulateGrad pass over
        # 85 MB, and ranges of the bucket can be all-reduced while the sweep continues: data_parallel.py)
        sink = graph.grad_sink
        direct = {}
        if sink is not None:
            for t, i in zip(graph.grad_params, graph._grad_slot):
                view = sink.view_of(t)
                if view is not None:
                    direct[i] = view
                    poff[i] = -1
            ptotal = 0
            for i in graph._grad_slot:
                if i not in direct:
                    poff[i] = ptotal
                    ptotal += (numel[i] + 3) // 4 * 4
        pflat = x0.new_empty(max(ptotal, 1))
        pbase = pflat.data_ptr()
        pgr = (c_void_p * n)()
        for i in graph._grad_slot:
            pgr[i] = direct[i].data_ptr() if i in direct else pbase + 4 * poff[i]
        written = (c_uint8 * n)()
        scratch_n = max(rows[v] * graph.values[v][0] for v in range(nv))
        scratch = x0.new_empty(max(scratch_n, 1))
        ptrs, tags = graph._param_arrays(None)
        marks = sink.marks() if sink is not None else None
        n_marks = len(marks[0]) if marks else 0
        mark_ops = (c_int32 * max(n_marks, 1))(*(marks[0] if marks else [0]))
        mark_evs = (c_void_p * max(n_marks, 1))(*(marks[1] if marks else [None]))
        _lib.check(_lib.lib.scn_graph_backward_marked(metadata._h, graph._c_ops, len(graph.ops), nv, ctx.vals, rows, ptrs,
                                                      tags, _lib.ptr(bn_save), grads, ext, pgr, written,
                                                      _lib.ptr(scratch), scratch_n, _lib.precision(), _lib.stream(),
                                                      n_marks, mark_ops, mark_evs))
        pg = []
        for t, i in zip(graph.grad_params, graph._grad_slot):
            if i in direct:
                pg.append(None)              # already in place (p.grad is the bucket view)
            else:
                pg.append(pflat[poff[i]:poff[i] + numel[i]].view_as(t) if written[i] else None)
        ctx.keep = None
        return (None, None, None, d_x0) + tuple(pg)
