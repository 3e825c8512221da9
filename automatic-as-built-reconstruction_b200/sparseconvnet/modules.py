"""The scn layers the sparse3d backbone is built from, with the reference's constructor
signatures, attribute / parameter names (state_dict compatible) and call semantics.

Reference files mirrored (all under SparseConvNet/sparseconvnet/): ioLayers.py:15-86,
submanifoldConvolution.py:14-113, convolution.py:13-126, deconvolution.py:13-155,
batchNormalization.py:14-187, networkInNetwork.py:14-88, sparseToDense.py:25-78,
sequential.py:9-61, tables.py:13-56, identity.py:10-15.  Each autograd Function calls the
pybind-compatible surface in SCN.py, which calls the CUDA library through the C ABI.
"""
import torch
from torch.autograd import Function
from torch.nn import Module, Parameter

import sparseconvnet
from . import SCN
from .tensor import SparseConvNetTensor


# ---- small helpers (utils.py:11-26) ------------------------------------------------------
def toLongTensor(dimension, x):
    if isinstance(x, torch.Tensor) and x.dtype == torch.int64 and not x.is_cuda:
        return x
    if isinstance(x, (list, tuple)):
        assert len(x) == dimension
        return torch.tensor(list(x), dtype=torch.int64)
    return torch.full((dimension,), int(x), dtype=torch.int64)


def optionalTensor(obj, name):
    return getattr(obj, name) if hasattr(obj, name) else torch.Tensor()


def optionalTensorReturn(t):
    return t if t.numel() else None


def Metadata(dim):
    """metadata.py:16-17"""
    return getattr(SCN, "Metadata_%d" % dim)()


def _like(input_tensor, features=None):
    out = SparseConvNetTensor()
    out.metadata = input_tensor.metadata
    out.spatial_size = input_tensor.spatial_size
    out.features = features
    return out


def _size_repr(t):
    v = t.tolist()
    return str(v[0]) if len(set(v)) == 1 else "(" + ",".join(str(i) for i in v) + ")"


def _conv_weight(filter_volume, groups, n_in, n_out):
    # He-style init of the reference: N(0, sqrt(2*groups / nIn / K))
    std = (2.0 * groups / n_in / filter_volume) ** 0.5
    return Parameter(torch.empty(filter_volume, groups, n_in // groups, n_out // groups).normal_(0, std))


# ---- InputLayer / OutputLayer ------------------------------------------------------------
class InputLayerFunction(Function):
    @staticmethod
    def forward(ctx, dimension, metadata, spatial_size, coords, input_features, batch_size, mode):
        ctx.metadata_ = metadata
        out = input_features.new_empty(0)
        prepared = coords.n_active if isinstance(coords, PreparedInput) else None
        SCN.InputLayer_updateOutput(metadata, spatial_size, None if prepared is not None else coords,
                                    input_features.contiguous(), out, batch_size, mode, prepared)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        grad_input = grad_output.new_empty(0)
        SCN.InputLayer_updateGradInput(ctx.metadata_, grad_input, grad_output.contiguous())
        return None, None, None, None, grad_input, None, None


class OutputLayerFunction(Function):
    @staticmethod
    def forward(ctx, dimension, metadata, input_features):
        ctx.metadata_ = metadata
        out = input_features.new_empty(0)
        SCN.OutputLayer_updateOutput(metadata, input_features.contiguous(), out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        grad_input = grad_output.new_empty(0)
        SCN.OutputLayer_updateGradInput(ctx.metadata_, grad_input, grad_output.contiguous())
        return None, None, grad_input


class PreparedInput(object):
    """The integer work of one batch, done ahead of time: a Metadata whose input grid (and, through
    FPN_Net.prepare, every downstream grid / rulebook) is already built, plus the CUDA event that
    marks its completion on the stream it was built on.  Pass it where coords would go:
    `net([prepared, features])`."""

    def __init__(self, metadata, spatial_size, n_active, n_points, event, coords):
        self.metadata, self.spatial_size, self.n_active, self.n_points = metadata, spatial_size, n_active, n_points
        self.event, self.coords = event, coords      # coords kept alive until the build has consumed them
        self.rulebooks_built = False


class InputLayer(Module):
    """(coords, features[, batch_size]) -> SparseConvNetTensor.

    coords: Long [N, dim] or [N, dim+1] (last column = sample index).  The reference always
    moves them to the CPU (ioLayers.py:60); here a CUDA coords tensor is used in place and a
    CPU one is uploaded once.  mode: 0 unique, 1/2 keep one point per voxel, 3 sum, 4 mean.
    """

    def __init__(self, dimension, spatial_size, mode=3):
        Module.__init__(self)
        self.dimension = dimension
        self.spatial_size = toLongTensor(dimension, spatial_size)
        self.mode = mode
        self.device = None

    def to(self, device):
        self.device = device
        return self

    def prepare(self, coords, batch_size=0):
        """Build the input hash grid for `coords` on the CURRENT stream and return a PreparedInput
        (B200 extension: the integer work depends on the coordinates only, so it can be overlapped
        with the previous step's feature kernels - see InputPrefetcher)."""
        m = Metadata(self.dimension)
        n_active = SCN.InputLayer_prepare(m, self.spatial_size, coords, batch_size, self.mode)
        ev = torch.cuda.Event()
        ev.record()
        return PreparedInput(m, self.spatial_size, n_active, coords.size(0), ev, coords)

    def forward(self, input):
        coords, feats = input[0], input[1]
        if self.device is not None:
            feats = feats.to(self.device)
        if isinstance(coords, PreparedInput):
            torch.cuda.current_stream().wait_event(coords.event)
            out = SparseConvNetTensor(metadata=coords.metadata, spatial_size=self.spatial_size)
            out.features = InputLayerFunction.apply(self.dimension, out.metadata, self.spatial_size, coords,
                                                    feats, 0, self.mode)
            out.rulebooks_built = coords.rulebooks_built
            return out
        out = SparseConvNetTensor(metadata=Metadata(self.dimension), spatial_size=self.spatial_size)
        out.features = InputLayerFunction.apply(self.dimension, out.metadata, self.spatial_size,
                                                coords.long(), feats,
                                                0 if len(input) == 2 else input[2], self.mode)
        return out


_loader_streams = {}


def loader_stream(device):
    """ONE high-priority side stream per device for the integer work that runs a batch ahead (InputPrefetcher,
    VoxelLoader): the library keeps per-stream workspaces, scan state and scheduling counters, so a fresh stream
    per loader would re-pay their first-use allocations (a device-synchronising cudaMalloc among them: measured
    180-280 ms for the first batch of a new loader) and never release them"""
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _loader_streams.get(key)
    if st is None:
        # high priority: the build is a few hundred tiny launches separated by count read-backs; each must slip in
        # beside the main stream's SM-filling kernels instead of queueing behind them
        st = _loader_streams[key] = torch.cuda.Stream(device=device, priority=-1)
    return st


class InputPrefetcher(object):
    """Runs `prepare_fn(coords)` (e.g. FPN_Net.prepare) for the NEXT batch on a side stream and a
    worker thread while the current batch's feature kernels run on the main stream - the role the
    DataLoader workers play for voxelisation in the reference (data3d/data.py:39-40).  The C calls
    release the GIL, and their count read-backs only block the worker.

        pf = scn.InputPrefetcher(net.prepare)
        pf.submit(coords0)
        for ...:
            prepared = pf.get(); pf.submit(next_coords)
            out = net([prepared, features])
    """

    def __init__(self, prepare_fn, device=None):
        import queue
        import threading
        self.prepare_fn = prepare_fn
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.stream = loader_stream(self.device)
        self.todo, self.done = queue.Queue(), queue.Queue()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        torch.cuda.set_device(self.device)
        while True:
            coords = self.todo.get()
            if coords is None:
                return
            try:
                with torch.cuda.stream(self.stream):
                    self.done.put(self.prepare_fn(coords))
            except Exception as e:  # noqa: BLE001  (re-raised in get())
                self.done.put(e)

    def submit(self, coords):
        self.todo.put(coords)

    def get(self):
        r = self.done.get()
        if isinstance(r, Exception):
            raise r
        return r

    def close(self):
        self.todo.put(None)


class OutputLayer(Module):
    def __init__(self, dimension):
        Module.__init__(self)
        self.dimension = dimension

    def forward(self, input):
        return OutputLayerFunction.apply(self.dimension, input.metadata, input.features)


class InputLayerInput(object):
    def __init__(self, coords, features):
        self.x = [coords, features]

    def __getitem__(self, n):
        return self.x[n]

    def __len__(self):
        return 2

    def cuda(self):
        self.x[1] = self.x[1].cuda()
        return self


# ---- convolutions ------------------------------------------------------------------------
class SubmanifoldConvolutionFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias, metadata, spatial_size, dimension, filter_size):
        ctx.metadata_ = metadata
        ctx.sizes = (spatial_size, filter_size)
        ctx.save_for_backward(input_features, weight, bias)
        out = input_features.new_empty(0)
        sparseconvnet.forward_pass_multiplyAdd_count += SCN.SubmanifoldConvolution_updateOutput(
            spatial_size, filter_size, metadata, input_features, out, weight, bias)
        sparseconvnet.forward_pass_hidden_states += out.nelement()
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, weight, bias = ctx.saved_tensors
        spatial_size, filter_size = ctx.sizes
        grad_input = grad_output.new_empty(0)
        grad_weight = torch.empty_like(weight)   # fully written by the kernel
        grad_bias = torch.zeros_like(bias)
        SCN.SubmanifoldConvolution_backward(spatial_size, filter_size, ctx.metadata_, input_features,
                                            grad_input, grad_output.contiguous(), weight,
                                            grad_weight, grad_bias)
        return grad_input, grad_weight, optionalTensorReturn(grad_bias), None, None, None, None


class SubmanifoldConvolution(Module):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        Module.__init__(self)
        self.dimension = dimension
        self.groups = groups
        self.nIn = nIn
        self.nOut = nOut
        self.filter_size = toLongTensor(dimension, filter_size)
        self.filter_volume = self.filter_size.prod().item()
        self.weight = _conv_weight(self.filter_volume, groups, nIn, nOut)
        if bias:
            self.bias = Parameter(torch.zeros(nOut))

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, \
            (self.nIn, self.nOut, input)
        return _like(input, SubmanifoldConvolutionFunction.apply(
            input.features, self.weight, optionalTensor(self, "bias"), input.metadata,
            input.spatial_size, self.dimension, self.filter_size))

    def __repr__(self):
        return "SubmanifoldConvolution %d->%d C%s" % (self.nIn, self.nOut, _size_repr(self.filter_size))

    def input_spatial_size(self, out_size):
        return out_size


class ValidConvolution(SubmanifoldConvolution):
    pass


class _StridedFunction(Function):
    """shared by Convolution and Deconvolution; `kind` picks the SCN entry points"""

    @staticmethod
    def forward(ctx, kind, input_features, weight, bias, metadata, in_size, out_size, dimension,
                filter_size, filter_stride):
        ctx.metadata_ = metadata
        ctx.kind = kind
        ctx.sizes = (in_size, out_size, filter_size, filter_stride)
        ctx.save_for_backward(input_features, weight, bias)
        out = input_features.new_empty(0)
        fwd = SCN.Convolution_updateOutput if kind == "conv" else SCN.Deconvolution_updateOutput
        sparseconvnet.forward_pass_multiplyAdd_count += fwd(
            in_size, out_size, filter_size, filter_stride, metadata, input_features, out, weight, bias)
        sparseconvnet.forward_pass_hidden_states += out.nelement()
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, weight, bias = ctx.saved_tensors
        in_size, out_size, filter_size, filter_stride = ctx.sizes
        grad_input = grad_output.new_empty(0)
        grad_weight = torch.empty_like(weight)
        grad_bias = torch.zeros_like(bias)
        bwd = SCN.Convolution_backward if ctx.kind == "conv" else SCN.Deconvolution_backward
        bwd(in_size, out_size, filter_size, filter_stride, ctx.metadata_, input_features, grad_input,
            grad_output.contiguous(), weight, grad_weight, grad_bias)
        return (None, grad_input, grad_weight, optionalTensorReturn(grad_bias), None, None, None,
                None, None, None)


class ConvolutionFunction(object):
    @staticmethod
    def apply(input_features, weight, bias, metadata, in_size, out_size, dimension, filter_size,
              filter_stride):
        return _StridedFunction.apply("conv", input_features, weight, bias, metadata, in_size,
                                      out_size, dimension, filter_size, filter_stride)


class DeconvolutionFunction(object):
    @staticmethod
    def apply(input_features, weight, bias, metadata, in_size, out_size, dimension, filter_size,
              filter_stride):
        return _StridedFunction.apply("deconv", input_features, weight, bias, metadata, in_size,
                                      out_size, dimension, filter_size, filter_stride)


class _StridedBase(Module):
    _name = ""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        Module.__init__(self)
        self.dimension = dimension
        self.groups = groups
        self.nIn = nIn
        self.nOut = nOut
        self.filter_size = toLongTensor(dimension, filter_size)
        self.filter_volume = self.filter_size.prod().item()
        self.filter_stride = toLongTensor(dimension, filter_stride)
        self.weight = _conv_weight(self.filter_volume, groups, nIn, nOut)
        if bias:
            self.bias = Parameter(torch.zeros(nOut))

    def __repr__(self):
        if len(set(self.filter_size.tolist())) == 1 and len(set(self.filter_stride.tolist())) == 1:
            s = "%d/%d" % (self.filter_size[0].item(), self.filter_stride[0].item())
        else:
            # the reference prints both tuples in full here, e.g. "C(1,1,8)/(1,1,1)" (convolution.py:57-64)
            s = "(%s)/(%s)" % (",".join(str(i) for i in self.filter_size.tolist()),
                               ",".join(str(i) for i in self.filter_stride.tolist()))
        return "%s %d->%d C%s" % (self._name, self.nIn, self.nOut, s)


class Convolution(_StridedBase):
    _name = "Convolution"

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        out_size = (input.spatial_size - self.filter_size) // self.filter_stride + 1
        assert ((out_size - 1) * self.filter_stride + self.filter_size == input.spatial_size).all(), \
            (input.spatial_size, out_size, self.filter_size, self.filter_stride)
        out = _like(input)
        out.spatial_size = out_size
        out.features = ConvolutionFunction.apply(
            input.features, self.weight, optionalTensor(self, "bias"), input.metadata,
            input.spatial_size, out_size, self.dimension, self.filter_size, self.filter_stride)
        return out

    def input_spatial_size(self, out_size):
        return (out_size - 1) * self.filter_stride + self.filter_size


class Deconvolution(_StridedBase):
    _name = "Deconvolution"

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        out_size = (input.spatial_size - 1) * self.filter_stride + self.filter_size
        out = _like(input)
        out.spatial_size = out_size
        out.features = DeconvolutionFunction.apply(
            input.features, self.weight, optionalTensor(self, "bias"), input.metadata,
            input.spatial_size, out_size, self.dimension, self.filter_size, self.filter_stride)
        return out

    def input_spatial_size(self, out_size):
        in_size = (out_size - self.filter_size) // self.filter_stride + 1
        assert ((in_size - 1) * self.filter_stride + self.filter_size == out_size).all()
        return in_size


# ---- NetworkInNetwork --------------------------------------------------------------------
class NetworkInNetworkFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias):
        ctx.save_for_backward(input_features, weight, bias)
        out = input_features.new_empty(0)
        sparseconvnet.forward_pass_multiplyAdd_count += SCN.NetworkInNetwork_updateOutput(
            input_features, out, weight, bias)
        sparseconvnet.forward_pass_hidden_states += out.nelement()
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input_features, weight, bias = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_input = grad_output.new_empty(0)
        grad_weight = torch.empty_like(weight)
        grad_bias = torch.zeros_like(bias)
        SCN.NetworkInNetwork_updateGradInput(grad_input, grad_output, weight)
        SCN.NetworkInNetwork_accGradParameters(input_features, grad_output, grad_weight, grad_bias)
        return grad_input, grad_weight, optionalTensorReturn(grad_bias)


class NetworkInNetwork(Module):
    def __init__(self, nIn, nOut, bias):
        Module.__init__(self)
        self.nIn = nIn
        self.nOut = nOut
        self.weight = Parameter(torch.empty(nIn, nOut).normal_(0, (2.0 / nIn) ** 0.5))
        if bias:
            self.bias = Parameter(torch.zeros(nOut))

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, \
            (self.nIn, input.features.shape)
        return _like(input, NetworkInNetworkFunction.apply(input.features, self.weight,
                                                           optionalTensor(self, "bias")))

    def __repr__(self):
        return "NetworkInNetwork%d->%d" % (self.nIn, self.nOut)

    def input_spatial_size(self, out_size):
        return out_size


# ---- BatchNormalization (+ ReLU / LeakyReLU) ---------------------------------------------
class BatchNormalizationFunction(Function):
    @staticmethod
    def forward(ctx, input_features, weight, bias, running_mean, running_var, eps, momentum, train,
                leakiness):
        ctx.train = train is True or train == 1
        ctx.leakiness = leakiness
        n_planes = running_mean.shape[0]
        out = input_features.new_empty(0)
        save_mean = input_features.new_empty(n_planes)
        save_invstd = input_features.new_empty(n_planes)
        SCN.BatchNormalization_updateOutput(input_features, out, save_mean, save_invstd, running_mean,
                                            running_var, weight, bias, eps, momentum, train, leakiness)
        ctx.save_for_backward(input_features, out, weight, bias, running_mean, running_var, save_mean,
                              save_invstd)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (input_features, output_features, weight, bias, running_mean, running_var, save_mean,
         save_invstd) = ctx.saved_tensors
        assert ctx.train
        grad_input = grad_output.new_empty(0)
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros_like(bias)
        SCN.BatchNormalization_backward(input_features, grad_input, output_features,
                                        grad_output.contiguous(), save_mean, save_invstd,
                                        running_mean, running_var, weight, bias, grad_weight,
                                        grad_bias, ctx.leakiness)
        return (grad_input, optionalTensorReturn(grad_weight), optionalTensorReturn(grad_bias), None,
                None, None, None, None, None)


class BatchNormalization(Module):
    """leakiness: 0 = ReLU, (0,1) = LeakyReLU, 1 = no activation.  `momentum` weights the OLD
    running value.  With track_running_stats=False, eval mode normalises with the statistics
    of the current batch (unbiased variance) - batchNormalization.py:51-56."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1,
                 track_running_stats=True):
        Module.__init__(self)
        self.nPlanes = nPlanes
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.leakiness = leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = Parameter(torch.ones(nPlanes))
            self.bias = Parameter(torch.zeros(nPlanes))
        self.track_running_stats = track_running_stats

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nPlanes, \
            (self.nPlanes, input.features.shape)
        # eval with track_running_stats=False (every shipped config): the reference computes features.mean(0) /
        # features.var(0) eagerly and hands them to the eval path (batchNormalization.py:51-56); here the
        # statistics kernel does it in the same launch pair as training (mode 2: batch statistics, unbiased
        # variance, running buffers untouched)
        mode = self.training if (self.training or self.track_running_stats) else 2
        return _like(input, BatchNormalizationFunction.apply(
            input.features, optionalTensor(self, "weight"), optionalTensor(self, "bias"), self.running_mean,
            self.running_var, self.eps, self.momentum, mode, self.leakiness))

    def input_spatial_size(self, out_size):
        return out_size

    def _repr(self, name, with_leak):
        s = "%s(%d,eps=%s,momentum=%s,affine=%s" % (name, self.nPlanes, self.eps, self.momentum,
                                                    self.affine)
        if with_leak:
            s += ",leakiness=" + str(self.leakiness)
        return s + ")"

    def __repr__(self):
        return self._repr("BatchNorm", self.leakiness > 0)


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, track_running_stats=True):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, 0, track_running_stats)

    def __repr__(self):
        return self._repr("BatchNormReLU", False)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333, track_running_stats=True):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, leakiness, track_running_stats)

    def __repr__(self):
        return self._repr("BatchNormLeakyReLU", True)


# ---- SparseToDense -----------------------------------------------------------------------
class SparseToDenseFunction(Function):
    @staticmethod
    def forward(ctx, input_features, metadata, spatial_size, dimension, n_planes):
        ctx.metadata_ = metadata
        ctx.spatial_size = spatial_size
        ctx.save_for_backward(input_features)
        out = input_features.new_empty(0)
        SCN.SparseToDense_updateOutput(spatial_size, metadata, input_features, out, n_planes)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (input_features,) = ctx.saved_tensors
        grad_input = grad_output.new_empty(0)
        SCN.SparseToDense_updateGradInput(ctx.spatial_size, ctx.metadata_, input_features, grad_input,
                                          grad_output.contiguous())
        return grad_input, None, None, None, None


class SparseToDense(Module):
    """SparseConvNetTensor -> dense [batch, nPlanes, X, Y, Z].  `dimension` is only stored: the
    kernel is chosen by the metadata's type, as in the reference (tools_3d_2d.py:25 passes 4)."""

    def __init__(self, dimension, nPlanes):
        Module.__init__(self)
        self.dimension = dimension
        self.nPlanes = nPlanes

    def forward(self, input):
        return SparseToDenseFunction.apply(input.features, input.metadata, input.spatial_size,
                                           self.dimension, self.nPlanes)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "SparseToDense(%d,%d)" % (self.dimension, self.nPlanes)


# ---- containers ----------------------------------------------------------------------------
class Sequential(torch.nn.Sequential):
    def input_spatial_size(self, out_size):
        for name in reversed(list(self._modules)):
            out_size = self._modules[name].input_spatial_size(out_size)
        return out_size

    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    append = add

    def insert(self, index, module):
        for i in range(len(self._modules), index, -1):
            self._modules[str(i)] = self._modules[str(i - 1)]
        self._modules[str(index)] = module


class CheckpointedSequential(Sequential):
    def forward(self, x):
        import torch.utils.checkpoint
        return torch.utils.checkpoint.checkpoint(lambda t: Sequential.forward(self, t), x)


class Identity(Module):
    def forward(self, input):
        return input

    def input_spatial_size(self, out_size):
        return out_size


class _Table(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def input_spatial_size(self, out_size):
        return out_size


class JoinTable(_Table):
    def forward(self, input):
        f0 = input[0].features
        return _like(input[0], torch.cat([i.features for i in input], 1) if f0.numel() else f0)


def _add_features(tensors):
    # not sum(): that starts from the int 0 and costs one extra full-size elementwise kernel
    out = tensors[0].features
    for t in tensors[1:]:
        out = out + t.features
    return out


class AddTable(_Table):
    def forward(self, input):
        return _like(input[0], _add_features(input))


class ConcatTable(_Table):
    def forward(self, input):
        return [module(input) for module in self._modules.values()]

    def input_spatial_size(self, out_size):
        return self._modules["0"].input_spatial_size(out_size)


def add_feature_planes(input):
    return _like(input[0], _add_features(input))


def concatenate_feature_planes(input):
    return _like(input[0], torch.cat([i.features for i in input], 1))
