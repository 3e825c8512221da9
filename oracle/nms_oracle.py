"""nms_oracle.py - CPU restatement of the rotated BEV IoU and the rotated NMS.
TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else.

  rotate_iou_eval:  second/core/non_max_suppression/nms_gpu.py:166-404 (corners, contained vertices, edge
                    intersections, angular sort, triangle-fan area) and :552-623 devRotateIoUEval criteria -1..2
  rotate_iou:       :626-703 rotate_iou_kernel_eval / rotate_iou_gpu_eval -> iou[n, k] = eval(query[k], boxes[n])
  rotate_nms:       second/pytorch/core/box_torch_ops.py:557-582 rotate_nms_3d (top pre_max_size by score, greedy,
                    post_max_size) -> second/core/non_max_suppression/nms_cpu.py:32-44 -> spconv 1.x
                    rotate_non_max_suppression_cpu (un-vendored; its published algorithm: boxes in score order, a kept
                    box suppresses every later box whose overlap IoU is >= thresh)
Pinned against tests/golden/rotate_iou.npz, which oracle/make_golden_nms.py produced by running the reference's own
numba device functions under numba's CUDA simulator.  Scalar float32 Python loops: small cases only."""
import math

import numpy as np

f32 = np.float32


def _corners(b):
    ca, sa = f32(math.cos(b[4])), f32(math.sin(b[4]))
    hx, hy = f32(b[2] / f32(2)), f32(b[3] / f32(2))
    px, py = [-hx, -hx, hx, hx], [-hy, hy, hy, -hy]
    c = np.zeros(8, f32)
    for i in range(4):
        c[2 * i] = ca * px[i] + sa * py[i] + b[0]
        c[2 * i + 1] = -sa * px[i] + ca * py[i] + b[1]
    return c


def _inside(x, y, c):
    ab0, ab1, ad0, ad1 = c[2] - c[0], c[3] - c[1], c[6] - c[0], c[7] - c[1]
    ap0, ap1 = x - c[0], y - c[1]
    abab, abap = ab0 * ab0 + ab1 * ab1, ab0 * ap0 + ab1 * ap1
    adad, adap = ad0 * ad0 + ad1 * ad1, ad0 * ap0 + ad1 * ap1
    return abab >= abap and abap >= 0 and adad >= adap and adap >= 0


def _cross(p, q, i, j):
    A, B = p[2 * i:2 * i + 2], p[2 * ((i + 1) % 4):2 * ((i + 1) % 4) + 2]
    C, D = q[2 * j:2 * j + 2], q[2 * ((j + 1) % 4):2 * ((j + 1) % 4) + 2]
    BA0, BA1, DA0, CA0, DA1, CA1 = B[0] - A[0], B[1] - A[1], D[0] - A[0], C[0] - A[0], D[1] - A[1], C[1] - A[1]
    acd = DA1 * CA0 > CA1 * DA0
    bcd = (D[1] - B[1]) * (C[0] - B[0]) > (C[1] - B[1]) * (D[0] - B[0])
    if acd == bcd:
        return None
    if (CA1 * BA0 > BA1 * CA0) == (DA1 * BA0 > BA1 * DA0):
        return None
    DC0, DC1 = D[0] - C[0], D[1] - C[1]
    ABBA, CDDC = A[0] * B[1] - B[0] * A[1], C[0] * D[1] - D[0] * C[1]
    DH = BA1 * DC0 - BA0 * DC1
    with np.errstate(all="ignore"):
        return f32((ABBA * DC0 - BA0 * CDDC) / DH), f32((ABBA * DC1 - BA1 * CDDC) / DH)


def intersection_area(b1, b2):
    c1, c2 = _corners(b1), _corners(b2)
    pts = []
    for i in range(4):
        if _inside(c1[2 * i], c1[2 * i + 1], c2):
            pts.append((c1[2 * i], c1[2 * i + 1]))
        if _inside(c2[2 * i], c2[2 * i + 1], c1):
            pts.append((c2[2 * i], c2[2 * i + 1]))
    for i in range(4):
        for j in range(4):
            t = _cross(c1, c2, i, j)
            if t is not None:
                pts.append(t)
    pts = pts[:8]                       # the reference's scratch holds 8 points
    n = len(pts)
    if n:
        cx = f32(sum((p[0] for p in pts), f32(0)) / f32(n))
        cy = f32(sum((p[1] for p in pts), f32(0)) / f32(n))
        key = []
        for p in pts:
            vx, vy = p[0] - cx, p[1] - cy
            with np.errstate(all="ignore"):
                d = f32(math.sqrt(vx * vx + vy * vy))
                vx, vy = f32(vx / d), f32(vy / d)
            key.append(f32(-2) - vx if vy < 0 else vx)
        pts, key = list(pts), list(key)
        for i in range(1, n):           # the reference's insertion sort (NaN keys compare false, as there)
            if key[i - 1] > key[i]:
                k, t = key[i], pts[i]
                j = i
                while j > 0 and key[j - 1] > k:
                    key[j], pts[j] = key[j - 1], pts[j - 1]
                    j -= 1
                key[j], pts[j] = k, t
    area = 0.0
    for i in range(n - 2):
        a, b, c = pts[0], pts[i + 1], pts[i + 2]
        area += abs(float(f32((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0]))) / 2.0)
    return f32(area)


def rotate_iou_eval(r1, r2, criterion=-1):
    """r1 = the query row, r2 = the box row"""
    r1, r2 = np.asarray(r1, f32), np.asarray(r2, f32)
    a1, a2 = r1[2] * r1[3], r2[2] * r2[3]
    inter = intersection_area(r1, r2)
    with np.errstate(all="ignore"):
        if criterion == -1:
            return f32(inter / (a1 + a2 - inter))
        if criterion == 0:
            return f32(inter / a1)
        if criterion == 1:
            return f32(inter / a2)
        if criterion == 2:
            if min(r2[2], r2[3]) / max(r2[2], r2[3]) < f32(0.25):
                return f32(inter / (a2 + max(f32(0), a1 * f32(0.5) - inter)))
            return f32(inter / (a1 + a2 - inter))
    return inter


def rotate_iou(boxes, query, criterion=-1):
    out = np.zeros((len(boxes), len(query)), f32)
    for n in range(len(boxes)):
        for k in range(len(query)):
            out[n, k] = rotate_iou_eval(query[k], boxes[n], criterion)
    return out


def rotate_nms(boxes, scores, iou_threshold, pre_max_size=None, post_max_size=None, iou=None):
    """indices kept, descending score.  boxes [n, 5] BEV rows; stable order for equal scores.
    iou: optional precomputed [n, n] matrix of the same boxes (indexing by original index)."""
    order = np.argsort(-np.asarray(scores, np.float64), kind="stable")
    if pre_max_size is not None and pre_max_size > 0:
        order = order[:pre_max_size]
    keep, dead = [], np.zeros(len(order), bool)
    for a in range(len(order)):
        if dead[a]:
            continue
        if post_max_size is not None and post_max_size > 0 and len(keep) >= post_max_size:
            break
        keep.append(int(order[a]))
        for b in range(a + 1, len(order)):
            if dead[b]:
                continue
            v = iou[order[a], order[b]] if iou is not None else rotate_iou_eval(boxes[order[a]], boxes[order[b]], -1)
            if v > 0 and v >= iou_threshold:
                dead[b] = True
    return np.asarray(keep, np.int64)
