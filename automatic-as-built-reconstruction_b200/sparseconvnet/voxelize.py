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


def voxelize_batch(buildings, scale, full_scale, matrix=None, xyz_feature=True, device=None, buffers=None):
    """The dataset's quantisation + collate for a whole batch on the GPU (reference:
    SUNCGDataset.__getitem__, data3d/suncg_utils/suncg_dataset.py:126-188 + trainMerge, data3d/data.py:25-37).

    buildings: list of float32 [n_i, C] tensors (columns 0-2 xyz in metres, then the feature columns), on the
    host (pinned for an asynchronous upload) or already on the device.  matrix: the dataset's 3x3 `m` (default
    eye(3) * scale, i.e. no zoom / flip / rotation augmentation).  Returns (coords int64 [N,4] CUDA with the
    building index in column 3, feats float32 [N,C] CUDA with feats[:, 0:3] = voxel-space position / scale when
    xyz_feature) - what `net([coords, feats])` takes.  One H2D copy of 4 C bytes per point (36 B for the 9-column
    SUNCG elements) replaces numpy's float64 passes and the 32 B/point int64 coordinate upload.
    buffers: optional dict reused across calls (grow-only staging / output tensors; the returned tensors are then
    views of it - VoxelLoader's ring)."""
    import numpy as np
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if len(buildings) == 0:
        raise RuntimeError("voxelize_batch: empty batch")
    C = buildings[0].size(1)
    for b in buildings:
        if not (b.dim() == 2 and b.size(1) == C and b.dtype == torch.float32):
            raise RuntimeError("voxelize_batch: every building must be a float32 [n, %d] tensor" % C)
    counts = [int(b.size(0)) for b in buildings]
    first = torch.tensor([0] + list(np.cumsum(counts)[:-1]), dtype=torch.int64)
    n = sum(counts)
    if buffers is not None:
        if buffers.get("cap", -1) < n or buffers.get("C") != C:
            cap = n + n // 8
            buffers.setdefault("retired", []).append([buffers.get(k) for k in ("pts", "coords", "feats")])
            buffers.update(cap=cap, C=C, pts=torch.empty(cap, C, dtype=torch.float32, device=device),
                           coords=torch.empty(cap, 4, dtype=torch.int64, device=device),
                           feats=torch.empty(cap, C, dtype=torch.float32, device=device))
        pts = buffers["pts"][:n]
    else:
        pts = torch.empty(n, C, dtype=torch.float32, device=device)
    off = 0
    for b, k in zip(buildings, counts):                       # one async copy per building into the batch buffer
        pts[off:off + k].copy_(b, non_blocking=True)
        off += k
    first_dev = first.to(device, non_blocking=True)
    m = np.eye(3) * float(scale) if matrix is None else np.asarray(matrix, dtype=np.float64).reshape(3, 3)
    from ctypes import c_double
    mat = (c_double * 9)(*m.reshape(-1).tolist())
    if buffers is not None:
        coords, feats = buffers["coords"][:n], buffers["feats"][:n]
    else:
        coords = torch.empty(n, 4, dtype=torch.int64, device=device)
        feats = torch.empty(n, C, dtype=torch.float32, device=device)
    kept = c_int64()
    check(lib.scn_voxelize_batch(ptr(pts), n, C, ptr(first_dev), len(buildings), mat, float(scale), i64x3(full_scale),
                                 1 if xyz_feature else 0, ptr(coords), ptr(feats), byref(kept), stream()))
    return coords[:kept.value], feats[:kept.value]


class VoxelLoader(object):
    """Batches of RAW buildings -> (PreparedInput, features) one batch ahead, on a side stream and a worker thread:
    upload of the float32 points, voxelisation (`voxelize_batch`) and the whole integer work of the batch
    (`prepare_fn`, e.g. FPN_Net.prepare: hash grids + rulebooks) overlap the previous batch's feature kernels - the
    role of the reference's DataLoader workers (data3d/data.py:39-40; `num_workers = 0` while its DEBUG flag is set)
    plus its numpy quantiser.

        loader = scn.VoxelLoader(batches, net.prepare, scale=50, full_scale=[4096, 4096, 512])
        for prepared, feats in loader:
            rpn_maps, roi_maps = net([prepared, feats])

    `batches`: an iterable of lists of float32 [n_i, C] tensors (pinned host memory recommended).  The returned
    feature tensor is a view of the loader's buffer ring: use it within the step it was handed out for (clone it to
    keep it longer)."""

    def __init__(self, batches, prepare_fn, scale, full_scale, matrix=None, xyz_feature=True, device=None, depth=1):
        import queue
        import threading
        self.batches, self.prepare_fn = batches, prepare_fn
        self.args = (scale, full_scale, matrix, xyz_feature)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        from .modules import loader_stream
        self.stream = loader_stream(self.device)
        # One batch waits in the queue, one is being built, one is in use by the consumer: a ring of depth + 2
        # buffer sets, reused without any allocation per batch (a fresh torch allocation per batch on the side
        # stream could not be recycled before the consumer's stream had passed it, and fell through to cudaMalloc -
        # a device-wide stall every few batches).  Set k is rewritten for batch i + ring only after the consumer has
        # asked for batch i + 1, i.e. after every use of batch i has been enqueued: the consumer records an event
        # then and the loader's stream waits for it.
        self.ring = max(1, depth) + 2
        self.sets = [dict() for _ in range(self.ring)]
        self.free_ev = [None] * self.ring
        self.taken = 0
        self.done = queue.Queue(maxsize=max(1, depth))
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        torch.cuda.set_device(self.device)
        try:
            for i, buildings in enumerate(self.batches):
                with torch.cuda.stream(self.stream):
                    k = i % self.ring
                    if self.free_ev[k] is not None:
                        self.stream.wait_event(self.free_ev[k])
                    scale, full_scale, matrix, xyz_feature = self.args
                    coords, feats = voxelize_batch(buildings, scale, full_scale, matrix, xyz_feature, self.device,
                                                   buffers=self.sets[k])
                    prepared = self.prepare_fn(coords)
                    ev = torch.cuda.Event()
                    ev.record()
                self.done.put((prepared, feats, ev))
            self.done.put(None)
        except Exception as e:  # noqa: BLE001  (re-raised by the consumer)
            self.done.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        if self.taken > 0:                                  # everything that uses the previous batch is enqueued
            fe = torch.cuda.Event()
            fe.record()
            self.free_ev[(self.taken - 1) % self.ring] = fe
        r = self.done.get()
        if r is None:
            raise StopIteration
        if isinstance(r, Exception):
            raise r
        prepared, feats, ev = r
        self.taken += 1
        torch.cuda.current_stream().wait_event(ev)          # the features were written on the loader's stream
        return prepared, feats
