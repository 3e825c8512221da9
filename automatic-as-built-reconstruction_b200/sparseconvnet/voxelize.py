"""GPU voxel quantisation front end (reference: numpy code in
data3d/suncg_utils/suncg_dataset.py:126-188 and the collate in data3d/data.py:25-37)."""
from ctypes import byref, c_int64

import torch

from ._lib import check, i64x3, lib, ptr, stream


def quantize_points(xyz, scale, full_scale, batch_idx=0):
    """xyz: CUDA float64 [n,3] metres.  Returns (coords int64 [m,4] CUDA, keep bool [n] CUDA):
    a = xyz*scale (fp64); a -= a.min(0); rows with a >= full_scale dropped; truncated to int64;
    batch column appended."""
    if not (xyz.is_cuda and xyz.dtype == torch.float64 and xyz.dim() == 2 and xyz.size(1) == 3):
        raise RuntimeError("quantize_points: xyz must be a CUDA float64 [n,3] tensor")
    xyz = xyz.contiguous()
    n = xyz.size(0)
    coords = torch.empty(n, 4, dtype=torch.int64, device=xyz.device)
    keep = torch.empty(n, dtype=torch.uint8, device=xyz.device)
    kept = c_int64()
    check(lib.scn_quantize_points(ptr(xyz), n, float(scale), i64x3(full_scale), int(batch_idx),
                                  ptr(coords), ptr(keep), byref(kept), stream()))
    return coords[:kept.value], keep.bool()
