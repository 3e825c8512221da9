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


def _pad4(n):
    return (n + 3) // 4 * 4


class GradBucket(object):
    """All gradients in one flat buffer; `p.grad` of every parameter is a view into it.

    GradBucket(params): plain bucket - autograd accumulates into the views, `allreduce_mean()` is ONE collective.

    GradBucket(params, module=net) with a module that runs through the layer-graph executor (FPN_Net): the bucket
    is laid out in the order the reverse sweep finishes the gradients and attached to the graph as its gradient
    SINK - `scn_graph_backward_marked` writes every weight / BN gradient straight into its bucket range (no
    AccumulateGrad pass) and records an event when each of `n_chunks` ranges is final (range ends at 1/2, 3/4, 7/8 ...
    of the bytes: the last range - the one collective nothing hides - is small).  `allreduce_mean()` then
    issues one NCCL all-reduce per range on a side stream behind those events, so the collective of the deep layers
    runs under the rest of the backward pass (DistributedDataParallel's overlapped buckets,
    tools/train_net_sparse3d.py:64-69); the main stream joins at the end.  Dead-branch parameters are never
    written and stay zero on every rank, so the collective's shape is the same everywhere.
    Direct writes OVERWRITE: call `zero()` (or nothing) between steps, not gradient accumulation over several
    backward passes - for that, build the bucket without `module`."""

    def __init__(self, params, module=None, n_chunks=6):
        params = [p for p in params if p.requires_grad]
        assert params, "no trainable parameters"
        graph = module._layer_graph() if (module is not None and hasattr(module, "_layer_graph")) else None
        self.chunks = []                      # (start, end, mark op) of the ranges the graph finishes, in order
        if graph is not None:
            last = graph.param_write_op()
            inside = sorted([p for p in params if id(p) in last], key=lambda p: -last[id(p)])   # stable
            outside = [p for p in params if id(p) not in last]
            params = inside + outside
            total_in = sum(_pad4(p.numel()) for p in inside)
            # range ends at 1/2, 3/4, 7/8, ... of the bytes: the reverse sweep finishes the deep layers - most of the
            # bytes - first and the small shallow weights last, so the LAST range (the only collective nothing hides)
            # carries ~3 % of the bucket instead of a quarter
            bounds = [total_in - (total_in >> k) for k in range(1, max(n_chunks, 1))] + [total_in]
            off, start, k = 0, 0, 0
            for p in inside:
                off += _pad4(p.numel())
                if off >= bounds[k] or p is inside[-1]:
                    self.chunks.append((start, off, last[id(p)]))
                    start = off
                    while k < len(bounds) - 1 and off >= bounds[k]:
                        k += 1
        self.params = params
        dev, dt = params[0].device, params[0].dtype
        n = sum(_pad4(p.numel()) for p in params)       # every range starts 16-byte aligned (vector stores)
        self.flat = torch.zeros(n, dtype=dt, device=dev)
        self._view, off = {}, 0
        for p in params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self._view[id(p)] = v
            off += _pad4(p.numel())
        self._events, self._fired, self._side = None, False, None
        if graph is not None and self.flat.is_cuda:
            import ctypes
            self._events = []
            for _ in self.chunks:
                e = ctypes.c_void_p()
                _lib.check(_lib.lib.scn_event_create(ctypes.byref(e)))
                self._events.append(e)
            module._grad_sink = self          # FPN_Net._layer_graph() hands it to the compiled graph
            graph.grad_sink = self

    # ---- the graph executor's side (graph.py GraphFunction.backward) -----------------------------------
    def view_of(self, p):
        """the bucket range of parameter p (None if it is not in the bucket); p.grad is re-attached if something
        (zero_grad(set_to_none=True)) dropped the view"""
        v = self._view.get(id(p))
        if v is not None and (p.grad is None or p.grad.data_ptr() != v.data_ptr()):
            p.grad = v
        return v

    def marks(self):
        if not self._events:
            return [], []
        self._fired = True
        return [c[2] for c in self.chunks], [e.value for e in self._events]

    # ---- the training loop's side ------------------------------------------------------------------------
    def check_views(self):
        """re-attach parameter gradients that no longer live in the bucket (optimizer.zero_grad() with
        set_to_none=True drops them: the next backward would then allocate fresh .grads and the all-reduce would
        see zeros).  Returns the number of re-attached parameters."""
        n = 0
        for p in self.params:
            v = self._view[id(p)]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
                n += 1
        return n

    def zero(self):
        self.flat.zero_()

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def allreduce_mean(self, group=None, async_op=False):
        """sum over ranks (NCCL on GPU tensors), then scale by 1/world with the library's kernel"""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        fired, self._fired = self._fired, False
        if world == 1:
            return None
        self.check_views()
        # NCCL averages inside the collective (no pass over the bucket afterwards); other backends: sum, then scale
        avg = self.flat.is_cuda and dist.get_backend(group) == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        if fired and self._events and not async_op:
            # overlapped: one collective per finished range, each behind its event, on a side stream
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.flat.device, priority=-1)
            works, end = [], 0
            for (a, b, _), ev in zip(self.chunks, self._events):
                _lib.check(_lib.lib.scn_stream_wait_event(self._side.cuda_stream, ev))
                with torch.cuda.stream(self._side):
                    works.append(dist.all_reduce(self.flat[a:b], op=op, group=group, async_op=True))
                end = b
            if end < self.flat.numel():       # parameters outside the graph: final once the main stream gets here
                works.append(dist.all_reduce(self.flat[end:], op=op, group=group, async_op=True))
            for w in works:
                w.wait()                      # the current stream waits for the collective (no host block)
            if not avg:
                self._scale(1.0 / world)
            return None
        work = dist.all_reduce(self.flat, op=op, group=group, async_op=async_op)
        if async_op:
            return _Pending(self, work, 1 if avg else world)
        if not avg:
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
        if self.world != 1:
            self.bucket._scale(1.0 / self.world)


def broadcast_parameters(module, src=0, group=None):
    """make every rank start from rank `src`'s weights (what DDP does at construction)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)
