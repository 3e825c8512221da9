"""sparse_3d_to_dense_2d (reference: SparseConvNet/sparseconvnet/tools_3d_2d.py:7-48): densify a
sparse map and crop it to the occupied extent - the ROI pooler's input."""
import sparseconvnet as scn


def sparse_3d_to_dense_2d(feat_s3d):
    loc = feat_s3d.get_spatial_locations()            # [x,y,z,batch] on the CPU
    x_size, y_size, z_size, _batch = (loc.max(0)[0] + 1).tolist()
    dense = scn.sparseToDense.SparseToDense(dimension=4, nPlanes=feat_s3d.features.shape[1])(feat_s3d)
    return dense[:, :, 0:x_size, 0:y_size, 0:z_size]   # [batch, C, x, y, z]
