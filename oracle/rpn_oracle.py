"""rpn_oracle.py - CPU restatement of the RPN's per-site work.  TEST INFRASTRUCTURE ONLY.

  rpn_head:      maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:108-131 (RPNHead.forward) with the input reshape of
                 RPNModule.forward (:184-186) - plain torch CPU ops, as the reference itself runs them
  grid_anchors:  maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:88-104
  example scope: .../anchor_generator_sparse3d.py:141-142, 174-185 (examples_bidx_2_sizes)
  base anchors:  .../anchor_generator_sparse3d.py:207-241 (generate_anchors_3d)
The reference modules import the compiled maskrcnn_benchmark `_C` extension (CUDA-only, not buildable here), so these
few lines are restated rather than imported; they are torch / numpy one-liners taken verbatim in meaning."""
import numpy as np
import torch
import torch.nn.functional as F


def rpn_head(features, conv_w, conv_b, cls_w, cls_b, box_w, box_b, A, S):
    """features [n, C] -> (logit [1, n, A, S], reg [1, n, A, 7 S])"""
    f = features.t().unsqueeze(0).unsqueeze(3)                 # rpn_sparse3d.py:184-186
    t = F.relu(F.conv2d(f, conv_w, conv_b))
    logit = F.conv2d(t, cls_w, cls_b).permute(0, 2, 1, 3)
    logit = logit.reshape(1, logit.shape[1], A, S)
    reg = F.conv2d(t, box_w, box_b).permute(0, 2, 1, 3)
    reg = reg.reshape(1, reg.shape[1], A, 7 * S)
    return logit, reg


def generate_anchors_3d(size, yaws, ratios, use_yaw):
    size = np.asarray(size, dtype=np.float32)
    out = []
    if use_yaw:
        for y in np.asarray(yaws, dtype=np.float32).reshape(-1, 1):
            out.append(np.concatenate([np.zeros(3), size, y]).reshape(1, -1))
    else:
        for r in np.asarray(ratios, dtype=np.float32):
            out.append(np.concatenate([np.zeros(3), size * r, np.array([0], np.float32)]).reshape(1, -1))
    return torch.from_numpy(np.concatenate(out, 0)).float()


def grid_anchors(locations, cell_anchors, voxel_scale, strides):
    anchors = []
    for base_anchors, location, stride in zip(cell_anchors, locations, strides):
        stride = torch.as_tensor(stride, dtype=torch.float32)
        c = (location[:, 0:3].float() + 0) / voxel_scale * stride.view(1, 3)
        c = torch.cat([c, torch.zeros(c.shape[0], 4)], 1).view(-1, 1, 7)
        anchors.append((c + base_anchors.view(1, -1, 7)).reshape(-1, 7))
    return anchors


def examples_bidx_2_sizes(examples_bidx, batch_size=None):
    batch_size = int(examples_bidx[-1]) + 1 if batch_size is None else batch_size
    s, out = 0, []
    for bi in range(batch_size):
        e = s + int(torch.sum(examples_bidx == bi))
        out.append([s, e])
        s = e
    return torch.tensor(out, dtype=torch.int64).view(-1, 2)
