"""roi_oracle.py - CPU restatement of the step right after the backbone:
SparseToDense -> crop -> ROIAlignRotated3D, forward and backward.

TEST INFRASTRUCTURE ONLY (tests/, tools/): never imported by the product package.

Restated, line by line, from
  maskrcnn_benchmark/layers/roi_align_rotated_3d.py:69-85      ROIAlignRotated3D.forward
  SparseConvNet/sparseconvnet/tools_3d_2d.py:7-26              sparse_3d_to_dense_2d (dense tensor cropped to
                                                               max active coordinate + 1 per axis)
  maskrcnn_benchmark/csrc/cuda/ROIAlignRotated3D_cuda.cu:16-87   bilinear_interpolate (8 corners)
  .../ROIAlignRotated3D_cuda.cu:90-178                          RoIAlignRotated3DForward
  .../ROIAlignRotated3D_cuda.cu:182-236, 238-355                bilinear_interpolate_gradient, ...BackwardFeature
The reference has NO CPU implementation of this operator (csrc/cpu holds ROIAlign_cpu / nms_cpu only) and its CUDA
build needs the GPU the build container lacks, so this restatement is NOT pinned against a run of the reference:
"parity unpinned" for row f1 (it is pinned against an independent dense fp64 trilinear evaluation in
tests/test_roi_oracle.py instead).  Arithmetic is float32 in the order of the CUDA source; cos / sin are the
platform's float routines, so agreement with any other float implementation is to rounding, not bit-exact.

Quirks kept: the forward bounds test reads `zsize > zsize` (always false, :28) - a sample above the volume is NOT
rejected in the forward pass but clamps to the top slice, while the backward pass (:190) rejects it; the dense
tensor's H axis is the sparse x axis and W the sparse y axis, while roi[1] / roi[4] ("w") run along W and
roi[2] / roi[5] ("h") along H."""
import numpy as np

F = np.float32


def dense_cropped(features, site_coords, batch_size):
    """sparse_3d_to_dense_2d: [B, C, X, Y, Z] with X,Y,Z = max active coordinate + 1 (over the whole batch)"""
    f = np.asarray(features, dtype=np.float32)
    c = np.asarray(site_coords, dtype=np.int64)
    ext = (c[:, :3].max(0) + 1) if len(c) else np.zeros(3, np.int64)
    out = np.zeros((batch_size, f.shape[1], int(ext[0]), int(ext[1]), int(ext[2])), np.float32)
    if len(c):
        out[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]] = f
    return out


def _sample_points(roi, spatial_scale, pooled, sampling_ratio):
    """float32 sample coordinates (y, x, z) [PH,PW,PZ,GH,GW,GZ] of one roi and the sample count per bin"""
    ph_n, pw_n, pz_n = pooled
    s = F(spatial_scale)
    cw, ch, cz = F(roi[1]) * s, F(roi[2]) * s, F(roi[3]) * s
    rw, rh, rz = F(roi[4]) * s, F(roi[5]) * s, F(roi[6]) * s
    theta = F(np.float64(F(roi[7])) * np.pi / 180.0)
    rw, rh, rz = max(rw, F(1.)), max(rh, F(1.)), max(rz, F(1.))
    bh, bw, bz = F(rh / F(ph_n)), F(rw / F(pw_n)), F(rz / F(pz_n))
    gh = sampling_ratio if sampling_ratio > 0 else int(np.ceil(F(rh / F(ph_n))))
    gw = sampling_ratio if sampling_ratio > 0 else int(np.ceil(F(rw / F(pw_n))))
    gz = sampling_ratio if sampling_ratio > 0 else int(np.ceil(F(rz / F(pz_n))))
    sh, sw, sz = F(-rh / F(2.0)), F(-rw / F(2.0)), F(-rz / F(2.0))
    ct, st = F(np.cos(theta)), F(np.sin(theta))

    def axis(start, n, b, g):
        p = np.arange(n, dtype=np.float32)[:, None]
        i = (np.arange(g, dtype=np.float32) + F(.5))[None, :]
        return (start + p * b + i * b / F(g)).astype(np.float32)          # [n, g]
    yy, xx, zz = axis(sh, ph_n, bh, gh), axis(sw, pw_n, bw, gw), axis(sz, pz_n, bz, gz)
    YY = yy[:, None, None, :, None, None]
    XX = xx[None, :, None, None, :, None]
    ZZ = zz[None, None, :, None, None, :]
    x = (XX * ct + YY * st + cw).astype(np.float32) + np.zeros_like(ZZ)
    y = (YY * ct - XX * st + ch).astype(np.float32) + np.zeros_like(ZZ)
    z = (ZZ + cz).astype(np.float32) + np.zeros_like(x)
    return y, x, z, gh * gw * gz


def _corners(y, x, z, H, W, Z, backward):
    """valid mask, the 8 (yi, xi, zi) corner index triples and their weights, in the order w1..w8 of the source"""
    if backward:
        valid = ~((y < -1.0) | (y > H) | (x < -1.0) | (x > W) | (z < -1.0) | (z > Z))
    else:
        valid = ~((y < -1.0) | (y > H) | (x < -1.0) | (x > W) | (z < -1.0))      # `zsize > zsize` never fires
    y, x, z = np.maximum(y, F(0)), np.maximum(x, F(0)), np.maximum(z, F(0))

    def lohi(v, n):
        lo = v.astype(np.int64)
        top = lo >= n - 1
        lo = np.where(top, n - 1, lo)
        hi = np.where(top, n - 1, lo + 1)
        v = np.where(top, lo.astype(np.float32), v)
        return lo, hi, (v - lo.astype(np.float32)).astype(np.float32)
    yl, yh, ly = lohi(y, H)
    xl, xh, lx = lohi(x, W)
    zl, zh, lz = lohi(z, Z)
    hy, hx, hz = F(1.) - ly, F(1.) - lx, F(1.) - lz
    idx = [(yl, xl, zl), (yl, xh, zl), (yh, xl, zl), (yh, xh, zl), (yl, xl, zh), (yl, xh, zh), (yh, xl, zh), (yh, xh, zh)]
    w = [hy * hx * hz, hy * lx * hz, ly * hx * hz, ly * lx * hz, hy * hx * lz, hy * lx * lz, ly * hx * lz, ly * lx * lz]
    return valid, idx, w


def roi_align_rotated_3d_forward(dense, rois, spatial_scale, pooled, sampling_ratio):
    """dense [B,C,H,W,Z] float32, rois [n,8] (batch, cw, ch, cz, w, h, z, theta_deg) -> [n,C,PH,PW,PZ]"""
    B, C, H, W, Z = dense.shape
    out = np.zeros((len(rois), C) + tuple(pooled), np.float32)
    if H == 0 or W == 0 or Z == 0:
        return out
    for n, roi in enumerate(np.asarray(rois, dtype=np.float32)):
        y, x, z, count = _sample_points(roi, spatial_scale, pooled, sampling_ratio)
        valid, idx, w = _corners(y, x, z, H, W, Z, backward=False)
        vol = dense[int(roi[0])]                                  # [C,H,W,Z]
        acc = np.zeros((C,) + y.shape, np.float32)
        for (yi, xi, zi), wi in zip(idx, w):
            acc += (wi * valid)[None] * vol[:, yi, xi, zi]
        out[n] = acc.sum(axis=(4, 5, 6)) / F(count)
    return out


def roi_align_rotated_3d_backward(grad_out, rois, spatial_scale, pooled, sampling_ratio, dense_shape):
    """gradient w.r.t. the dense input, [B,C,H,W,Z]"""
    B, C, H, W, Z = dense_shape
    g = np.zeros(dense_shape, np.float32)
    if H == 0 or W == 0 or Z == 0:
        return g
    for n, roi in enumerate(np.asarray(rois, dtype=np.float32)):
        y, x, z, count = _sample_points(roi, spatial_scale, pooled, sampling_ratio)
        valid, idx, w = _corners(y, x, z, H, W, Z, backward=True)
        top = grad_out[n][:, :, :, :, None, None, None]            # [C,PH,PW,PZ,1,1,1]
        vol = g[int(roi[0])]
        for (yi, xi, zi), wi in zip(idx, w):
            contrib = (top * ((wi * valid) / F(count))[None]).reshape(C, -1)
            flat = ((yi * W + xi) * Z + zi).reshape(-1)
            np.add.at(vol.reshape(C, -1), (slice(None), flat), contrib)
    return g


def sparse_roi_align(features, site_coords, batch_size, rois, spatial_scale, pooled, sampling_ratio, grad_out=None):
    """the reference module on a sparse map: returns out (and, with grad_out, the gradient at the active sites =
    SparseToDense backward of the dense gradient, sparseToDense.py / CPU/SparseToDense.cpp:66-87)"""
    c = np.asarray(site_coords, dtype=np.int64)
    dense = dense_cropped(features, c, batch_size)
    out = roi_align_rotated_3d_forward(dense, rois, spatial_scale, pooled, sampling_ratio)
    if grad_out is None:
        return out
    gd = roi_align_rotated_3d_backward(np.asarray(grad_out, np.float32), rois, spatial_scale, pooled, sampling_ratio,
                                       dense.shape)
    return out, gd[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]]
