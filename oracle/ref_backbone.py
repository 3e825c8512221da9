"""ref_backbone.py - drive the COMPILED REFERENCE (oracle/_ref/SCN_ref.so) through the sparse3d
backbone graph on the CPU.  TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs; never imported by the product package.

The reference's Python layer files do not travel to the GPU box (/root/reference is absent
there), so this file restates - functionally, on a state_dict - what they do:
  sparseconvnet/ioLayers.py:51-64,163-218        InputLayer (mode 4)
  sparseconvnet/submanifoldConvolution.py:30-113 SubmanifoldConvolution fwd/bwd
  sparseconvnet/convolution.py:31-126            Convolution fwd/bwd (with the `//` fix)
  sparseconvnet/deconvolution.py:31-155          Deconvolution fwd/bwd
  sparseconvnet/batchNormalization.py:46-187     BatchNorm(Leaky)ReLU fwd/bwd, eval semantics
  sparseconvnet/fpn_net.py:95-265                the FPN_Net graph (reps=1, residual blocks)
Every arithmetic op is executed by the reference's own C++ (SCN_ref.*), unmodified.
tests/test_oracle_vs_ref.py checks this driver against the reference's real fpn_net.py when
/root/reference is importable (in the build container).
"""
import importlib.util
import os

import torch
from torch.autograd import Function

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    path = os.path.join(_HERE, "_ref", name + ".so")
    if not os.path.exists(path):
        raise ImportError("%s missing: run `python oracle/build_ref.py` in the build container" % path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_SCN = None
_DUMP = None


def scn_ref():
    global _SCN
    if _SCN is None:
        _SCN = _load("SCN_ref")
    return _SCN


def scn_refdump():
    global _DUMP
    if _DUMP is None:
        _DUMP = _load("SCN_refdump")
    return _DUMP


def available():
    return all(os.path.exists(os.path.join(_HERE, "_ref", n + ".so")) for n in ("SCN_ref", "SCN_refdump"))


def L(v):
    return torch.tensor([int(i) for i in v], dtype=torch.int64)


_E = torch.Tensor  # empty optional tensor


class _Subm(Function):
    @staticmethod
    def forward(ctx, x, w, md, ss, fs):
        ctx.md, ctx.ss, ctx.fs = md, ss, fs
        ctx.save_for_backward(x, w)
        y = x.new()
        scn_ref().SubmanifoldConvolution_updateOutput(ss, fs, md, x, y, w, _E())
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw = dy.new(), torch.zeros_like(w)
        scn_ref().SubmanifoldConvolution_backward(ctx.ss, ctx.fs, ctx.md, x, dx, dy.contiguous(), w, dw, _E())
        return dx, dw, None, None, None


class _Strided(Function):
    @staticmethod
    def forward(ctx, x, w, md, in_ss, out_ss, fs, st, deconv):
        ctx.a = (md, in_ss, out_ss, fs, st, deconv)
        ctx.save_for_backward(x, w)
        y = x.new()
        f = scn_ref().Deconvolution_updateOutput if deconv else scn_ref().Convolution_updateOutput
        f(in_ss, out_ss, fs, st, md, x, y, w, _E())
        return y

    @staticmethod
    def backward(ctx, dy):
        md, in_ss, out_ss, fs, st, deconv = ctx.a
        x, w = ctx.saved_tensors
        dx, dw = dy.new(), torch.zeros_like(w)
        f = scn_ref().Deconvolution_backward if deconv else scn_ref().Convolution_backward
        f(in_ss, out_ss, fs, st, md, x, dx, dy.contiguous(), w, dw, _E())
        return dx, dw, None, None, None, None, None, None


class _BN(Function):
    @staticmethod
    def forward(ctx, x, w, b, rm, rv, eps, momentum, train, leak):
        ctx.leak, ctx.train = leak, train
        y = x.new()
        sm, si = x.new(rm.shape[0]), x.new(rm.shape[0])
        scn_ref().BatchNormalization_updateOutput(x, y, sm, si, rm, rv, w, b, eps, momentum, train, leak)
        ctx.save_for_backward(x, y, w, b, rm, rv, sm, si)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, w, b, rm, rv, sm, si = ctx.saved_tensors
        assert ctx.train
        dx, dw, db = dy.new(), torch.zeros_like(w), torch.zeros_like(b)
        # the reference overwrites grad_output in place (CPU/BatchNormalization.cpp:79-82): clone
        scn_ref().BatchNormalization_backward(x, dx, y, dy.contiguous().clone(), sm, si, rm, rv, w, b, dw,
                                              db, ctx.leak)
        return dx, dw, db, None, None, None, None, None, None


class _Input(Function):
    @staticmethod
    def forward(ctx, feats, md, ss, coords, batch_size, mode):
        ctx.md = md
        y = feats.new()
        scn_ref().InputLayer_updateOutput(md, ss, coords, feats.contiguous(), y, batch_size, mode)
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = dy.new()
        scn_ref().InputLayer_updateGradInput(ctx.md, dx, dy.contiguous())
        return dx, None, None, None, None, None


class RefTensor(object):
    def __init__(self, features, md, ss):
        self.features, self.md, self.ss = features, md, ss

    def locations(self):
        return self.md.getSpatialLocations(self.ss)


class RefBackbone(object):
    """FPN_Net graph (fpn_net.py:95-265) on reference ops, parameters taken from a state_dict with
    the FPN_Net key names.  Parameters become leaf tensors; after backward() their .grad is set."""

    def __init__(self, state_dict, full_scale=(4096, 4096, 512), n_planes=(32, 64, 64, 128, 128, 128, 256, 256, 256),
                 down_kernels=None, down_strides=None, fpn_scales_from_top=(4, 3, 2, 1),
                 roi_scales_from_top=(4, 3), rpn_3d_2d_selector=(1, 2, 3, 4, 5, 6),
                 rpn_map_sizes=((256, 256, 32), (128, 128, 16), (64, 64, 8), (32, 32, 4)), leakiness=0,
                 bn_momentum=0.95, track_running_stats=False, eps=1e-4):
        self.p = {k: v.detach().clone().float().cpu() for k, v in state_dict.items()}
        for k, v in self.p.items():
            if "running_" not in k:
                v.requires_grad_(True)
        self.full_scale = L(full_scale)
        self.n_scales = len(n_planes)
        self.kernels = down_kernels or [[2, 2, 2]] * (self.n_scales - 1)
        self.strides = down_strides or [[2, 2, 2]] * (self.n_scales - 1)
        self.fpn_scales_from_top = list(fpn_scales_from_top)
        self.roi_scales_from_top = list(roi_scales_from_top)
        self.selector = list(rpn_3d_2d_selector)
        self.rpn_map_sizes = [list(s) for s in rpn_map_sizes]
        self.leak, self.momentum, self.track, self.eps = leakiness, bn_momentum, track_running_stats, eps
        self.training = True

    # ---- layers --------------------------------------------------------------------------
    def bn(self, key, t):
        p = self.p
        rm, rv = p[key + ".running_mean"], p[key + ".running_var"]
        if not (self.training or self.track):          # batchNormalization.py:51-56
            rm, rv = t.features.detach().mean(0), t.features.detach().var(0)
        y = _BN.apply(t.features, p[key + ".weight"], p[key + ".bias"], rm, rv, self.eps, self.momentum,
                      self.training, self.leak)
        return RefTensor(y, t.md, t.ss)

    def subm(self, key, t, fs):
        return RefTensor(_Subm.apply(t.features, self.p[key + ".weight"], t.md, t.ss, L([fs] * 3)), t.md, t.ss)

    def conv(self, key, t, fs, st):
        fs, st = L(fs), L(st)
        out_ss = (t.ss - fs) // st + 1                   # convolution.py:35-38
        assert ((out_ss - 1) * st + fs == t.ss).all()
        return RefTensor(_Strided.apply(t.features, self.p[key + ".weight"], t.md, t.ss, out_ss, fs, st, False),
                         t.md, out_ss)

    def deconv(self, key, t, fs, st):
        fs, st = L(fs), L(st)
        out_ss = (t.ss - 1) * st + fs                    # deconvolution.py:35-36
        return RefTensor(_Strided.apply(t.features, self.p[key + ".weight"], t.md, t.ss, out_ss, fs, st, True),
                         t.md, out_ss)

    def block(self, key, t):
        """ConcatTable(Identity, Sequential(BN, Subm3, BN, Subm3)) -> AddTable (fpn_net.py:60-69)"""
        h = self.bn(key + ".1.0", t)
        h = self.subm(key + ".1.1", h, 3)
        h = self.bn(key + ".1.2", h)
        h = self.subm(key + ".1.3", h, 3)
        return RefTensor(t.features + h.features, t.md, t.ss)

    # ---- graph ---------------------------------------------------------------------------
    def input_layer(self, coords, feats, mode=4):
        md = scn_ref().Metadata_3()
        y = _Input.apply(feats, md, self.full_scale, coords.cpu().long(), 0, mode)
        return RefTensor(y, md, self.full_scale)

    def forward(self, coords, feats):
        t = self.input_layer(coords, feats)
        t = self.subm("layers_in.1", t, 3)
        downs = []
        for k in range(self.n_scales):
            if k == 0:
                t = self.block("m_downs.0.0", t)
            else:
                t = self.bn("m_downs.%d.0.0" % k, t)
                t = self.conv("m_downs.%d.0.1" % k, t, self.kernels[k - 1], self.strides[k - 1])
                t = self.block("m_downs.%d.1" % k, t)
            downs.append(t)
        t = self.subm("m_shortcuts.%d" % (self.n_scales - 1), t, 1)
        ups = [t]
        for k in range(self.n_scales - 1):
            j = self.n_scales - 2 - k
            t = self.bn("m_ups.%d.0" % k, t)
            t = self.deconv("m_ups.%d.1" % k, t, self.kernels[j], self.strides[j])
            s = self.subm("m_shortcuts.%d" % j, downs[j], 1)
            t = RefTensor(t.features + s.features, t.md, t.ss)
            ups.append(self.subm("m_mergeds.%d" % k, t, 3))   # `net` itself stays un-merged (fpn_net.py:186-196)
        maps3d = [ups[i] for i in self.fpn_scales_from_top]
        maps2d = [self.conv("convs_pro2d.%d" % i, m, [1, 1, self.rpn_map_sizes[i][2]], [1, 1, 1])
                  for i, m in enumerate(maps3d)]
        both = maps3d + maps2d
        return [both[i] for i in self.selector], [ups[i] for i in self.roi_scales_from_top]

    def grads(self):
        return {k: v.grad for k, v in self.p.items() if v.requires_grad and v.grad is not None}


def backbone_loss(rpn_maps, roi_maps):
    """sum of squared features of the 6 rpn + 2 roi outputs"""
    return sum((m.features ** 2).sum() for m in list(rpn_maps) + list(roi_maps))
