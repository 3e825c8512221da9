"""Rotated IoU / rotated NMS.  CPU: oracle/nms_oracle.py against the golden values the
reference's own numba device functions produced (oracle/make_golden_nms.py).  GPU: csrc/nms.cu through the C ABI
against the golden values and the oracle.  IoU is float32 geometry: the bound is 2e-5 absolute (measured: sin / cos
and division rounding differ between the simulator's libm and the device by a few ulp of values <= 1); keep lists
are compared exactly on cases whose IoUs are not within that bound of the threshold."""
import os

import numpy as np
import pytest
import torch

import nms_oracle as NO

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rotate_iou.npz"))
TOL = 2e-5


def _close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    same_nan = np.isnan(a) == np.isnan(b)
    return bool(same_nan.all()) and float(np.nanmax(np.abs(a - b), initial=0.0)) <= TOL


@pytest.mark.parametrize("crit", [-1, 0, 1, 2])
def test_oracle_iou_matches_reference_golden(crit):
    got = NO.rotate_iou(GOLD["boxes"], GOLD["query"], crit)
    assert _close(got, GOLD["iou_c%d" % crit])


def test_oracle_self_iou_and_plain():
    assert _close(NO.rotate_iou(GOLD["boxes"], GOLD["boxes"], -1), GOLD["iou_self"])
    assert _close(NO.rotate_iou(GOLD["boxes"], GOLD["query"], -1), GOLD["iou_plain"])


def _nms_case(n, seed):
    rng = np.random.RandomState(seed)
    c = rng.rand(n, 2) * 8
    d = np.stack([rng.rand(n) * 3 + 0.2, rng.rand(n) * 0.5 + 0.05], 1)
    yaw = rng.choice([0, np.pi / 2, 0.4, -1.0], n) + rng.randn(n) * 0.05
    return np.concatenate([c, d, yaw[:, None]], 1).astype(np.float32), rng.rand(n).astype(np.float32)


def test_oracle_nms_properties():
    b, s = _nms_case(120, 3)
    iou = NO.rotate_iou(b, b, -1)
    keep = NO.rotate_nms(b, s, 0.1, iou=iou.T)
    assert (np.diff(s[keep]) <= 0).all()                                  # descending score
    for i, a in enumerate(keep):                                          # kept boxes do not suppress one another
        for c in keep[i + 1:]:
            assert not (iou[c, a] > 0 and iou[c, a] >= 0.1)
    dropped = sorted(set(range(len(b))) - set(keep.tolist()))
    for j in dropped:                                                     # every dropped box has a better keeper
        assert any(s[a] >= s[j] and iou[j, a] >= 0.1 for a in keep)
    top = np.argsort(-s, kind="stable")[:50]                              # pre / post limits
    assert np.array_equal(NO.rotate_nms(b, s, 0.1, pre_max_size=50, post_max_size=7, iou=iou.T),
                          top[NO.rotate_nms(b[top], s[top], 0.1)[:7]])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("crit", [-1, 0, 1, 2])
def test_gpu_iou_matches_reference_golden(crit):
    import sparseconvnet as scn
    got = scn.rotate_iou_gpu_eval(torch.from_numpy(GOLD["boxes"]).cuda(), torch.from_numpy(GOLD["query"]).cuda(), crit)
    assert got.is_cuda and _close(got.cpu().numpy(), GOLD["iou_c%d" % crit])
    if crit == -1:
        self_iou = scn.rotate_iou_gpu_eval(torch.from_numpy(GOLD["boxes"]).cuda(), torch.from_numpy(GOLD["boxes"]).cuda())
        assert _close(self_iou.cpu().numpy(), GOLD["iou_self"])


@pytest.mark.gpu
def test_gpu_iou_empty_and_3d_wrapper():
    import sparseconvnet as scn
    e = torch.zeros(0, 5).cuda()
    b = torch.from_numpy(GOLD["boxes"]).cuda()
    assert scn.rotate_iou_gpu_eval(e, b).shape == (0, 48) and scn.rotate_iou_gpu_eval(b, e).shape == (48, 0)
    b7 = torch.zeros(48, 7).cuda()
    b7[:, [0, 1, 3, 4, 6]] = b
    b7[:, 5] = 2.5
    assert _close(scn.boxes_iou_3d(b7, b7, flag="rpn_post").cpu().numpy(), GOLD["iou_self"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,thresh,pre,post", [(120, 0.1, None, None), (700, 0.1, 500, 100), (700, 0.3, 2000, 500),
                                               (65, 0.5, None, 3), (1, 0.1, 2000, 500), (3000, 0.1, 2000, 500)])
def test_gpu_nms_matches_oracle(n, thresh, pre, post):
    import sparseconvnet as scn
    b, s = _nms_case(n, n)
    bt, st = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
    iou = scn.rotate_iou_gpu_eval(bt, bt).cpu().numpy()         # iou[n, k] = eval(query k, box n): kept = query
    want = NO.rotate_nms(b, s, thresh, pre, post, iou=iou.T)
    b7 = torch.zeros(n, 7).cuda()
    b7[:, [0, 1, 3, 4, 6]] = bt
    got = scn.rotate_nms_3d(b7, st, pre_max_size=pre, post_max_size=post, iou_threshold=thresh, flag="rpn_post")
    assert got.is_cuda and got.dtype == torch.int64
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(scn.rotate_nms(bt, st, pre, post, thresh).cpu().numpy(), want)


@pytest.mark.gpu
def test_gpu_nms_small_oracle_geometry_and_empty():
    """the greedy pass with the oracle's own IoU (no device matrix in the loop)"""
    import sparseconvnet as scn
    b, s = _nms_case(60, 11)
    want = NO.rotate_nms(b, s, 0.15)
    got = scn.rotate_nms(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda(), None, None, 0.15)
    assert np.array_equal(got.cpu().numpy(), want)
    assert scn.rotate_nms(torch.zeros(0, 5).cuda(), torch.zeros(0).cuda()).numel() == 0
