"""Rotated BEV IoU and rotated NMS on the device (csrc/nms.cu).

Reference call path: maskrcnn_benchmark/structures/boxlist_ops_3d.py:13-70 boxlist_nms_3d ->
second/pytorch/core/box_torch_ops.py:557-582 rotate_nms_3d -> second/core/non_max_suppression/nms_cpu.py:32-44
rotate_nms_3d_cc -> utils3d/rotate_nms_3d_torch.py:22-100 boxes_iou_3d ->
second/core/non_max_suppression/nms_gpu.py:667-703 rotate_iou_gpu_eval.  The reference copies the boxes to the
host, the IoU matrix back, and walks it greedily on the CPU; here boxes and scores stay on the device and only the
number of kept boxes is read back.  Same names, arguments and results (indices into the input, descending score)."""
import ctypes

import torch

from ._lib import check, lib, ptr, require_cuda_f32, stream


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """iou[n, k] of boxes [N, 5] against query_boxes [K, 5], rows (x, y, size_x, size_y, yaw); CUDA tensors in and
    out (nms_gpu.py:667-703 takes and returns host arrays)"""
    boxes = require_cuda_f32(boxes, "boxes")
    query_boxes = require_cuda_f32(query_boxes, "query boxes")
    assert boxes.dim() == 2 and boxes.size(1) == 5 and query_boxes.dim() == 2 and query_boxes.size(1) == 5
    iou = boxes.new_zeros(boxes.size(0), query_boxes.size(0))
    check(lib.scn_rotate_iou(ptr(boxes), boxes.size(0), ptr(query_boxes), query_boxes.size(0), int(criterion),
                             ptr(iou), stream()))
    return iou


def boxes_iou_3d(targets_bbox3d, anchors_bbox3d, aug_thickness=None, criterion=-1, only_xy=True, flag=""):
    """rotate_nms_3d_torch.py:22-100 with its DEBUG switch as shipped (only_xy forced on: the BEV IoU is the
    result); rows (x, y, z, size_x, size_y, size_z, yaw)"""
    if aug_thickness is None:
        aug_thickness = {"target_Y": 0.0, "target_Z": 0.0, "anchor_Y": 0.0, "anchor_Z": 0.0}
    t = targets_bbox3d.detach().clone().float()
    a = anchors_bbox3d.detach().clone().float()
    t[:, 3].clamp_(min=aug_thickness["target_Y"])
    a[:, 3].clamp_(min=aug_thickness["anchor_Y"])
    return rotate_iou_gpu_eval(t[:, [0, 1, 3, 4, 6]].contiguous(), a[:, [0, 1, 3, 4, 6]].contiguous(), criterion)


def _nms(bev, scores, pre_max_size, post_max_size, iou_threshold):
    bev = require_cuda_f32(bev, "boxes")
    scores = require_cuda_f32(scores, "scores")
    n = bev.size(0)
    if n == 0:
        return torch.zeros([0], dtype=torch.int64, device=bev.device)
    pre = int(pre_max_size) if pre_max_size is not None else 0
    post = int(post_max_size) if post_max_size is not None else 0
    m = min(n, pre) if pre > 0 else n
    keep = torch.empty(min(m, post) if post > 0 else m, dtype=torch.int64, device=bev.device)
    n_keep = ctypes.c_int64(0)
    check(lib.scn_rotate_nms(ptr(bev), ptr(scores), n, float(iou_threshold), pre, post, ptr(keep),
                             ctypes.byref(n_keep), stream()))
    return keep[:n_keep.value]


def rotate_nms_3d(rbboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5, flag=""):
    """box_torch_ops.py:557-582: rbboxes [n, 7] (x, y, z, size_x, size_y, size_z, yaw)"""
    assert rbboxes.dim() == 2 and rbboxes.size(1) == 7
    return _nms(rbboxes[:, [0, 1, 3, 4, 6]].contiguous(), scores, pre_max_size, post_max_size, iou_threshold)


def rotate_nms(rbboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5):
    """box_torch_ops.py:527-555: rbboxes [n, 5] BEV rows"""
    assert rbboxes.dim() == 2 and rbboxes.size(1) == 5
    return _nms(rbboxes, scores, pre_max_size, post_max_size, iou_threshold)
