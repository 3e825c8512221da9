"""GPU parity of the sparse ROIAlignRotated3D (csrc/roi.cu) through the C ABI against
oracle/roi_oracle.py (SparseToDense -> crop -> the restated ROIAlignRotated3D_cuda.cu kernels).
Floating point: tolerance 1e-5 * max|reference| (float32 cos / sin and fused multiply-adds differ in the last
place between implementations; the operator is continuous in the sample position except on the volume boundary)."""
import numpy as np
import pytest
import torch

import roi_oracle as R

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _map(scn, n=3000, ss=(64, 48, 16), ext=(60, 44, 14), C=128, batch=2, seed=0):
    rng = np.random.RandomState(seed)
    c = np.concatenate([np.concatenate([(rng.rand(n, 3) * np.array(ext)).astype(np.int64), np.full((n, 1), b)], 1)
                        for b in range(batch)])
    torch.manual_seed(seed)
    feats = torch.randn(len(c), C)
    leaf = feats.cuda().requires_grad_(True)
    t = scn.InputLayer(3, list(ss), 4)([torch.from_numpy(c), leaf])
    return t, leaf


def _rois(n, ext, batch, seed, scale):
    rng = np.random.RandomState(seed)
    r = np.zeros((n, 8), np.float32)
    r[:, 0] = rng.randint(0, batch, n)
    r[:, 1] = rng.rand(n) * ext[1] / scale           # center_w runs along the sparse y axis
    r[:, 2] = rng.rand(n) * ext[0] / scale           # center_h along x
    r[:, 3] = rng.rand(n) * ext[2] / scale
    r[:, 4:7] = (rng.rand(n, 3) * np.array([20, 14, 8]) + 0.2) / scale
    r[:, 7] = rng.rand(n) * 360 - 180
    r[: n // 8, 1:4] += 30 / scale                   # some boxes partly / fully outside the volume
    r[n // 8: n // 6, 4:7] = 0.01                     # malformed tiny boxes
    return r


def _check(scn, C, pooled, sampling, scale, n_rois=24, seed=0):
    t, leaf = _map(scn, C=C, seed=seed)
    rois = _rois(n_rois, (60, 44, 14), 2, seed + 1, scale)
    pool = scn.ROIAlignRotated3D(pooled, scale, sampling)
    out = pool(t, torch.from_numpy(rois).cuda())
    assert list(out.shape) == [n_rois, C] + list(pooled)
    g = torch.randn_like(out)
    (dfeat,) = torch.autograd.grad(out, t.features, g)
    loc = t.get_spatial_locations().numpy()
    want, dwant = R.sparse_roi_align(t.features.detach().cpu().numpy(), loc, 2, rois, scale, pooled, sampling,
                                     grad_out=g.cpu().numpy())
    assert np.abs(out.detach().cpu().numpy() - want).max() <= TOL * np.abs(want).max()
    assert np.abs(dfeat.cpu().numpy() - dwant).max() <= TOL * max(np.abs(dwant).max(), 1e-30)


@pytest.mark.parametrize("C,pooled,sampling,scale", [(128, (5, 11, 4), 2, 0.25), (128, (7, 7, 7), 2, 1.0),
                                                     (32, (3, 4, 2), 0, 0.5), (7, (2, 2, 2), 1, 1.0),
                                                     (640, (2, 3, 2), 2, 0.5), (256, (1, 1, 1), 3, 0.125)])
def test_sparse_roi_align_matches_dense_oracle(C, pooled, sampling, scale):
    import sparseconvnet as scn
    _check(scn, C, pooled, sampling, scale)


def test_no_rois_and_module_surface():
    import sparseconvnet as scn
    t, _ = _map(scn, C=16)
    pool = scn.ROIAlignRotated3D((5, 11, 4), 0.25, 2)
    assert repr(pool) == "ROIAlignRotated3D(output_size=(5, 11, 4), spatial_scale=0.25, sampling_ratio=2)"
    out = pool(t, torch.zeros(0, 8).cuda())
    assert list(out.shape) == [0, 16, 5, 11, 4]
    with pytest.raises(RuntimeError):
        pool(t, torch.zeros(3, 5).cuda())
    with pytest.raises(RuntimeError):
        pool(t, torch.zeros(3, 8))          # no CPU path


def test_equals_the_dense_route_of_this_library():
    """the reference's own route - sparse_3d_to_dense_2d, then sampling the dense tensor - evaluated with this
    library's SparseToDense and the oracle's dense kernel restatement"""
    import sparseconvnet as scn
    t, _ = _map(scn, C=64, seed=3)
    rois = _rois(16, (60, 44, 14), 2, 9, 0.5)
    dense = scn.tools_3d_2d.sparse_3d_to_dense_2d(t).detach().cpu().numpy()
    want = R.roi_align_rotated_3d_forward(dense, rois, 0.5, (4, 4, 4), 2)
    out = scn.ROIAlignRotated3D((4, 4, 4), 0.5, 2)(t, torch.from_numpy(rois).cuda())
    assert np.abs(out.detach().cpu().numpy() - want).max() <= TOL * np.abs(want).max()
