"""ROIAlignRotated3D on a sparse map (reference: maskrcnn_benchmark/layers/roi_align_rotated_3d.py:11-98).

Same module surface - `ROIAlignRotated3D(output_size, spatial_scale, sampling_ratio)(input_s3d, rois_3d)`, the
functional `roi_align_rotated_3d`, the same roi layout [batch, center_w, center_h, center_z, width, height, zsize,
theta_deg] and `__repr__` - but the reference's first step, `sparse_3d_to_dense_2d(input_s3d)` (a dense
[B,C,X,Y,Z] tensor: 1.07 GB for a [1,128,256,256,32] map, zero-filled and scattered every call), is gone: the
kernel looks the trilinear corners up in the sparse map's hash grid (csrc/roi.cu).  Results are those of the
dense path: an inactive corner contributes 0, and the feature gradient is what SparseToDense's backward keeps."""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import check, i64x3, lib, ptr, require_cuda_f32, stream


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3, "output_size must be (pooled_height, pooled_width, pooled_zsize)"
        return tuple(int(i) for i in v)
    return (int(v),) * 3


class _ROIAlignRotated3DSparse(Function):
    @staticmethod
    def forward(ctx, features, metadata, spatial_size, roi, output_size, spatial_scale, sampling_ratio):
        x = require_cuda_f32(features, "input features")
        r = require_cuda_f32(roi, "rois")
        if r.dim() != 2 or r.size(1) != 8:
            raise RuntimeError("rois must be [n, 8] (batch, center_w, center_h, center_z, width, height, zsize, theta)")
        pooled = (ctypes.c_int64 * 3)(*output_size)
        out = x.new_empty(r.size(0), x.size(1) if x.dim() == 2 else 0, *output_size)
        if x.dim() == 2 and out.numel():
            check(lib.scn_roi_align_rotated_3d_forward(metadata._h, i64x3(spatial_size), ptr(x), x.size(1), ptr(r),
                                                       r.size(0), float(spatial_scale), pooled, int(sampling_ratio),
                                                       ptr(out), stream()))
        ctx.metadata_, ctx.spatial_size = metadata, spatial_size
        ctx.args = (output_size, float(spatial_scale), int(sampling_ratio), tuple(x.shape))
        ctx.save_for_backward(r)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (r,) = ctx.saved_tensors
        output_size, spatial_scale, sampling_ratio, shape = ctx.args
        g = require_cuda_f32(grad_output, "grad_output")
        d = g.new_empty(shape)
        if len(shape) == 2 and d.numel():
            pooled = (ctypes.c_int64 * 3)(*output_size)
            check(lib.scn_roi_align_rotated_3d_backward(ctx.metadata_._h, i64x3(ctx.spatial_size), ptr(g), shape[1],
                                                        ptr(r), r.size(0), spatial_scale, pooled, sampling_ratio,
                                                        ptr(d), stream()))
        return d, None, None, None, None, None, None


def roi_align_rotated_3d(input_s3d, roi, output_size, spatial_scale, sampling_ratio):
    """functional form on a SparseConvNetTensor (the reference's takes the densified tensor)"""
    return _ROIAlignRotated3DSparse.apply(input_s3d.features, input_s3d.metadata, input_s3d.spatial_size, roi,
                                          _triple(output_size), spatial_scale, sampling_ratio)


class ROIAlignRotated3D(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        """output_size: (pooled_height, pooled_width, pooled_zsize); spatial_scale: map size / original size;
        sampling_ratio: samples per bin and axis (<= 0: ceil(roi extent / pooled extent))"""
        super(ROIAlignRotated3D, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input_s3d, rois_3d):
        """input_s3d: sparse 3d tensor; rois_3d: [n, 8] boxes, xyz order as in input_s3d, yaw in degrees,
        anti-clockwise positive.  Returns [n, C, pooled_height, pooled_width, pooled_zsize]."""
        return roi_align_rotated_3d(input_s3d, rois_3d, self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio)
