"""RPN head and anchor generation on the device.

Reference: maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:84-131 (RPNHead) and
maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:88-146 (AnchorGenerator.grid_anchors / forward).

`RPNHead` keeps the reference's parameters - `conv`, `cls_logits`, `bbox_pred` as nn.Conv2d with 1x1 kernels, same
names, shapes and init, so state_dicts load unchanged - and its outputs (lists of [1, n, A, S] logits and
[1, n, A, 7 S] regressions).  It takes the levels either as the reference does ([1, C, n, 1] tensors made by
`features.t().unsqueeze(0).unsqueeze(3)`) or, without that transpose, as SparseConvNetTensors / [n, C] tensors; the
three 1x1 convolutions run as row-major GEMMs of the library (csrc/rpn.cu).  `grid_anchors` builds the anchors
from the sparse maps' device coordinates: no get_spatial_locations() host copy, no host -> device upload."""
import ctypes

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function

from . import _lib
from ._lib import check, i64x3, lib, ptr, require_cuda_f32, stream


class _RPNHeadFunction(Function):
    @staticmethod
    def forward(ctx, x, wc, bc, wl, bl, wr, br):
        x = require_cuda_f32(x, "RPN head features")
        n, C = x.shape
        ws = [require_cuda_f32(w, "RPN head weight") for w in (wc, wl, wr)]
        bs = [require_cuda_f32(b, "RPN head bias") for b in (bc, bl, br)]
        n_cls, n_box = wl.size(0), wr.size(0)
        hidden, logits, reg = x.new_empty(n, C), x.new_empty(n, n_cls), x.new_empty(n, n_box)
        check(lib.scn_rpn_head_forward(ptr(x), n, C, ptr(ws[0]), ptr(bs[0]), ptr(ws[1]), ptr(bs[1]), n_cls, ptr(ws[2]),
                                       ptr(bs[2]), n_box, ptr(hidden), ptr(logits), ptr(reg), _lib.precision(), stream()))
        ctx.save_for_backward(x, hidden, *ws)
        return logits, reg

    @staticmethod
    def backward(ctx, d_logits, d_reg):
        x, hidden, wc, wl, wr = ctx.saved_tensors
        n, C = x.shape
        n_cls, n_box = wl.size(0), wr.size(0)
        d_logits = require_cuda_f32(d_logits if d_logits is not None else x.new_zeros(n, n_cls), "grad")
        d_reg = require_cuda_f32(d_reg if d_reg is not None else x.new_zeros(n, n_box), "grad")
        d_hidden = torch.empty_like(hidden)
        d_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dwc, dwl, dwr = torch.empty_like(wc), torch.empty_like(wl), torch.empty_like(wr)
        dbc, dbl, dbr = x.new_empty(C), x.new_empty(n_cls), x.new_empty(n_box)
        check(lib.scn_rpn_head_backward(ptr(x), ptr(hidden), n, C, ptr(wc), ptr(wl), n_cls, ptr(wr), n_box, ptr(d_logits),
                                        ptr(d_reg), ptr(d_hidden), ptr(d_x), ptr(dwc), ptr(dbc), ptr(dwl), ptr(dbl),
                                        ptr(dwr), ptr(dbr), _lib.precision(), stream()))
        return d_x, dwc, dbc, dwl, dbl, dwr, dbr


class RPNHead(nn.Module):
    def __init__(self, in_channels, num_anchors_per_location, seperate_rpn=1):
        """in_channels: width of the FPN maps (cfg.SPARSE3D.nPlaneMap); num_anchors_per_location: yaws per site;
        seperate_rpn: int(len(cfg.MODEL.SEPARATE_CLASSES) * cfg.MODEL.SEPARATE_RPN) + 1 (rpn_sparse3d.py:95)"""
        super(RPNHead, self).__init__()
        self.num_anchors_per_location = num_anchors_per_location
        self.seperate_rpn = seperate_rpn
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.cls_logits = nn.Conv2d(in_channels, num_anchors_per_location * seperate_rpn, kernel_size=1, stride=1)
        self.bbox_pred = nn.Conv2d(in_channels, num_anchors_per_location * 7 * seperate_rpn, kernel_size=1, stride=1)
        for layer in (self.conv, self.cls_logits, self.bbox_pred):
            torch.nn.init.normal_(layer.weight, std=0.01)
            torch.nn.init.constant_(layer.bias, 0)

    def forward(self, x):
        logits, bbox_reg = [], []
        A, S = self.num_anchors_per_location, self.seperate_rpn
        for feature in x:
            f = getattr(feature, "features", feature)
            if f.dim() == 4:                       # the reference's [1, C, n, 1]
                f = f[0, :, :, 0].t()
            logit, reg = _RPNHeadFunction.apply(
                f, self.conv.weight.view(self.conv.out_channels, -1), self.conv.bias,
                self.cls_logits.weight.view(A * S, -1), self.cls_logits.bias,
                self.bbox_pred.weight.view(7 * A * S, -1), self.bbox_pred.bias)
            logits.append(logit.view(1, -1, A, S))
            bbox_reg.append(reg.view(1, -1, A, 7 * S))
        return logits, bbox_reg


def grid_anchors(feature_maps_sparse, cell_anchors, voxel_scale, strides, with_scope=True):
    """AnchorGenerator.grid_anchors + the per-sample scopes of AnchorGenerator.forward, on the device.
    feature_maps_sparse: the RPN levels (SparseConvNetTensors); cell_anchors: per level float [A, 7] base anchors
    (generate_anchors_3d); strides: per level 3 floats.  Returns (anchors: list of CUDA float [n_l * A, 7],
    scopes: list of CUDA int64 [batch, 2] row ranges (times A) of every sample, or None)."""
    anchors, scopes = [], []
    for fm, base, stride in zip(feature_maps_sparse, cell_anchors, strides):
        m, ss = fm.metadata, fm.spatial_size
        n = m.getNActive(ss)
        base = require_cuda_f32(base.to("cuda") if not base.is_cuda else base, "cell anchors").view(-1, 7)
        A = base.size(0)
        out = base.new_empty(max(n, 0) * A, 7)
        batch = m.getBatchSize()
        scope = torch.empty(batch, 2, dtype=torch.int64, device=base.device) if with_scope else None
        st = (ctypes.c_float * 3)(*[float(v) for v in stride])
        check(lib.scn_grid_anchors(m._h, i64x3(ss), ptr(base), A, float(voxel_scale), st, ptr(out), ptr(scope), batch,
                                   stream()))
        anchors.append(out)
        scopes.append(scope)
    return anchors, (scopes if with_scope else None)
