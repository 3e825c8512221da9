"""Data parallelism for the sparse3d backbone: one process per GPU, whole buildings per rank.

Reference behaviour: tools/train_net_sparse3d.py:64-69 wraps the
model in DistributedDataParallel(broadcast_buffers=False) - one bucketed gradient all-reduce per
step, BN statistics per GPU, no collective inside the backbone; data3d/data.py:39-40 has no
DistributedSampler, so partitioning buildings by rank is done here.  Inference
(engine/inference_3d.py:17-75) shards whole buildings per rank with no collective.

B200 design: all gradients live in ONE flat fp32 buffer (`GradBucket`); parameter `.grad`s are
views into it, so the weight-gradient kernels' results land in the bucket and the step's single
NCCL all-reduce (NVLink5 / NVSwitch, ~85 MB for the backbone) needs no packing copy.  Parameters
that receive no gradient (the dead FPN branches, layers_out, linear - fpn_net.py:198-203) stay
zero in the bucket on every rank, which keeps the collective shape identical across ranks.
"""
import torch
import torch.distributed as dist

from . import _lib


def shard_indices(n_items, rank, world_size):
    """buildings of rank r: r, r+W, r+2W, ... (whole-building sharding, no intra-building split)"""
    return list(range(rank, n_items, world_size))


class GradBucket(object):
    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def allreduce_mean(self, group=None, async_op=False):
        """sum over ranks (NCCL on GPU tensors), then scale by 1/world with the library's kernel"""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return _Pending(self, work, world)
        self._scale(1.0 / world)
        return None

    def _scale(self, alpha):
        if self.flat.is_cuda:
            _lib.check(_lib.lib.scn_scale_inplace(_lib.ptr(self.flat), alpha, self.flat.numel(), _lib.stream()))
        else:  # host-side logic tests (gloo)
            self.flat.mul_(alpha)


class _Pending(object):
    def __init__(self, bucket, work, world):
        self.bucket, self.work, self.world = bucket, work, world

    def wait(self):
        self.work.wait()
        self.bucket._scale(1.0 / self.world)


def broadcast_parameters(module, src=0, group=None):
    """make every rank start from rank `src`'s weights (what DDP does at construction)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)
