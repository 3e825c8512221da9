"""oracle/roi_oracle.py (restatement of ROIAlignRotated3D_cuda.cu) against an INDEPENDENT
evaluation: scipy's trilinear map_coordinates at sample points recomputed here in float64 - the reference has no
CPU implementation and no fixtures for this operator ("parity unpinned", see the oracle's header)."""
import numpy as np
from scipy.ndimage import map_coordinates

import roi_oracle as R


def _independent(dense, roi, scale, pooled, s):
    b, cw, ch, cz, w, h, z, th = [float(v) for v in roi]
    cw, ch, cz, w, h, z = [v * scale for v in (cw, ch, cz, w, h, z)]
    w, h, z = max(w, 1.0), max(h, 1.0), max(z, 1.0)
    t = th * np.pi / 180.0
    out = np.zeros((dense.shape[1],) + tuple(pooled))
    for ph in range(pooled[0]):
        for pw in range(pooled[1]):
            for pz in range(pooled[2]):
                pts = []
                for iy in range(s):
                    yy = -h / 2 + ph * h / pooled[0] + (iy + .5) * h / pooled[0] / s
                    for ix in range(s):
                        xx = -w / 2 + pw * w / pooled[1] + (ix + .5) * w / pooled[1] / s
                        for iz in range(s):
                            zz = -z / 2 + pz * z / pooled[2] + (iz + .5) * z / pooled[2] / s
                            pts.append((yy * np.cos(t) - xx * np.sin(t) + ch, xx * np.cos(t) + yy * np.sin(t) + cw, zz + cz))
                pts = np.array(pts).T
                for c in range(dense.shape[1]):
                    out[c, ph, pw, pz] = map_coordinates(dense[int(b), c].astype(np.float64), pts, order=1).mean()
    return out


def test_forward_matches_independent_trilinear_on_interior_rois():
    rng = np.random.RandomState(0)
    dense = rng.randn(2, 3, 20, 24, 10).astype(np.float32)
    rois = np.array([[0, 24, 20, 10, 12, 10, 6, 0], [1, 20, 22, 9, 10, 14, 5, 30], [0, 26, 18, 11, 9, 9, 4, -75]], np.float32)
    got = R.roi_align_rotated_3d_forward(dense, rois, 0.5, (3, 4, 2), 2)
    for n, roi in enumerate(rois):
        want = _independent(dense, roi, 0.5, (3, 4, 2), 2)
        assert np.abs(got[n] - want).max() <= 1e-5 * np.abs(want).max()


def test_backward_is_the_adjoint_of_forward():
    rng = np.random.RandomState(1)
    dense = rng.randn(1, 4, 16, 16, 8).astype(np.float32)
    rois = np.array([[0, 16, 15, 8, 10, 8, 6, 20], [0, 14, 17, 7, 7, 12, 5, 100]], np.float32)
    out = R.roi_align_rotated_3d_forward(dense, rois, 0.5, (2, 3, 2), 2)
    g = rng.randn(*out.shape).astype(np.float32)
    gd = R.roi_align_rotated_3d_backward(g, rois, 0.5, (2, 3, 2), 2, dense.shape)
    assert abs(float((out.astype(np.float64) * g).sum()) - float((dense.astype(np.float64) * gd).sum())) <= 1e-4 * float(np.abs(out * g).sum())


def test_quirks_and_edges():
    dense = np.ones((1, 1, 4, 4, 4), np.float32)
    # a sample above the volume: kept (clamped to the top slice) by the forward test `zsize > zsize`, dropped by
    # the backward one; a sample beyond y > height is dropped by both
    hi = np.array([[0, 2, 2, 9, 2, 2, 2, 0]], np.float32)
    assert R.roi_align_rotated_3d_forward(dense, hi, 1.0, (1, 1, 1), 1)[0, 0, 0, 0, 0] == 1.0
    assert R.roi_align_rotated_3d_backward(np.ones((1, 1, 1, 1, 1), np.float32), hi, 1.0, (1, 1, 1), 1, dense.shape).sum() == 0
    out = np.array([[0, 2, 9, 2, 2, 2, 2, 0]], np.float32)
    assert R.roi_align_rotated_3d_forward(dense, out, 1.0, (1, 1, 1), 1)[0, 0, 0, 0, 0] == 0.0
    # malformed (tiny) rois are forced to 1 x 1 x 1; adaptive sampling (ratio 0) takes ceil(extent / pooled)
    tiny = np.array([[0, 2, 2, 2, 0.01, 0.01, 0.01, 0]], np.float32)
    assert abs(R.roi_align_rotated_3d_forward(dense, tiny, 1.0, (2, 2, 2), 0)[0, 0].mean() - 1.0) < 1e-6
    assert R.roi_align_rotated_3d_forward(dense, np.zeros((0, 8), np.float32), 1.0, (2, 2, 2), 2).shape == (0, 1, 2, 2, 2)
    # sparse wrapper: cropped extent = max active coordinate + 1
    f = np.arange(6, dtype=np.float32).reshape(3, 2)
    c = np.array([[0, 0, 0, 0], [2, 1, 0, 0], [1, 3, 2, 1]])
    assert R.dense_cropped(f, c, 2).shape == (2, 2, 3, 4, 3)
