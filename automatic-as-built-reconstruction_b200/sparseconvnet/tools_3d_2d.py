"""sparse_3d_to_dense_2d (reference: SparseConvNet/sparseconvnet/tools_3d_2d.py:7-48): densify a
sparse map and crop it to the occupied extent - the ROI pooler's input.

The reference copies every location to the host for the extent (`get_spatial_locations().max(0)`), densifies ALL of
[X, Y, Z] (1.07 GB zero-filled for a [1,128,256,256,32] ROI map) and slices the occupied corner out.  Here the extent
is one 16-byte read-back and the scatter goes straight into the cropped volume - the same values as the reference's
slice (tests/test_gpu_parity.py::test_sparse_3d_to_dense_2d_cropped), contiguous instead of a view."""
import torch
from torch.autograd import Function

from . import SCN


class _CroppedDense(Function):
    @staticmethod
    def forward(ctx, input_features, metadata, spatial_size, extent):
        ctx.metadata_, ctx.spatial_size, ctx.extent = metadata, spatial_size, extent
        ctx.save_for_backward(input_features)
        out = input_features.new_empty(0)
        SCN.SparseToDense_cropped_updateOutput(spatial_size, extent, metadata, input_features, out,
                                               input_features.shape[1])
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (input_features,) = ctx.saved_tensors
        grad_input = grad_output.new_empty(0)
        SCN.SparseToDense_cropped_updateGradInput(ctx.spatial_size, ctx.extent, ctx.metadata_, input_features,
                                                  grad_input, grad_output.contiguous())
        return grad_input, None, None, None


def sparse_3d_to_dense_2d(feat_s3d):
    x_size, y_size, z_size, _batch = SCN.grid_extent(feat_s3d.metadata, feat_s3d.spatial_size)
    return _CroppedDense.apply(feat_s3d.features, feat_s3d.metadata, feat_s3d.spatial_size,
                               [x_size, y_size, z_size])                    # [batch, C, x, y, z]
