"""`sparseconvnet.SCN` - the extension-module surface of the reference, re-hosted on libscn_b200.

The reference builds `sparseconvnet.SCN` with pybind11 (SparseConvNet/sparseconvnet/SCN/
pybind.cpp:34-235) and every layer file calls `sparseconvnet.SCN.<Op>_updateOutput/_backward`.
This module keeps exactly those names, argument orders and in-place `resize_` semantics, so the
reference's own Python layer files could import it unchanged, and forwards each call through the
C ABI (include/scn_b200.h) with raw device pointers and torch's current CUDA stream.

Feature tensors must be CUDA float32 (sparseconvnet_cuda.cpp dispatches on `is_cuda()`; the
CPU branch does not exist here).  Only dimension 3 (`Metadata_3`) is implemented - the only
one the sparse3d backbone instantiates.
"""
import ctypes
from ctypes import byref, c_double, c_int64, c_void_p

import torch

from . import _lib
from ._lib import check, i64x3, lib, ptr, require_cuda_f32, stream


def n_rulebook_bits():  # pybind.cpp:234
    return lib.scn_n_rulebook_bits()


class Metadata_3(object):
    """Opaque device-resident Metadata<3> (reference class: Metadata/Metadata.h:44-163)."""

    dimension = 3

    def __init__(self):
        h = c_void_p()
        check(lib.scn_metadata_create(3, byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:          # `lib` is None while the interpreter shuts down
            lib.scn_metadata_destroy(h)

    # ---- reference methods used on the path -------------------------------------------
    def clear(self):
        check(lib.scn_metadata_clear(self._h, stream()))

    def getNActive(self, spatial_size):
        n = c_int64()
        check(lib.scn_get_nactive(self._h, i64x3(spatial_size), byref(n)))
        return n.value

    def getBatchSize(self):
        n = c_int64()
        check(lib.scn_get_batch_size(self._h, byref(n)))
        return n.value

    def getSpatialLocations(self, spatial_size):
        """CPU int64 [nActive,4] rows (x,y,z,batch) - Metadata.cpp:149-168"""
        n = self.getNActive(spatial_size)
        if n < 0:
            raise RuntimeError("Metadata: no grid at spatial size %s" % (list(spatial_size),))
        out = torch.zeros(n, 4, dtype=torch.int64)
        check(lib.scn_get_spatial_locations(self._h, i64x3(spatial_size), c_void_p(out.data_ptr()),
                                            stream()))
        return out

    def getSpatialLocationsDevice(self, spatial_size):
        """same rows as int32 on the GPU, no host synchronisation (B200 extension)"""
        n = self.getNActive(spatial_size)
        out = torch.empty(max(n, 0), 4, dtype=torch.int32, device="cuda")
        check(lib.scn_get_spatial_locations_device(self._h, i64x3(spatial_size), ptr(out), stream()))
        return out

    # ---- eager rulebook construction (B200 extension) ---------------------------------------
    # The reference builds rulebooks lazily inside the first layer that needs them
    # (Metadata.cpp:430-512); every build reads counts back to the host.  Building the whole
    # network's rulebooks up front keeps those read-backs out of the feature-kernel stream.
    def prepareSubmanifoldRuleBook(self, spatial_size, filter_size):
        check(lib.scn_submanifold_rulebook_prepare(self._h, i64x3(spatial_size), i64x3(filter_size),
                                                   stream(), None))

    def prepareRuleBook(self, in_size, out_size, filter_size, filter_stride):
        n_out = c_int64()
        check(lib.scn_conv_rulebook_prepare(self._h, i64x3(in_size), i64x3(out_size), i64x3(filter_size),
                                            i64x3(filter_stride), stream(), byref(n_out), None))
        return n_out.value

    # ---- rulebook export (parity tests; reference: Metadata.h:35 RuleBook) ----------------
    def inputLayerRuleBook(self):
        hd = (c_int64 * 4)()
        check(lib.scn_input_rulebook_header(self._h, hd, stream()))
        mode, max_active, n_in, n_out = list(hd)
        table = torch.zeros(n_out, 1 + max_active, dtype=torch.int32)
        if mode != 0 and n_out:
            check(lib.scn_input_rulebook_copy(self._h, c_void_p(table.data_ptr()), stream()))
        return [mode, max_active, n_in, n_out], table

    def getSubmanifoldRuleBook(self, spatial_size, filter_size):
        fs = i64x3(filter_size)
        K = int(fs[0] * fs[1] * fs[2])
        counts = (c_int64 * K)()
        check(lib.scn_submanifold_rulebook_prepare(self._h, i64x3(spatial_size), fs, stream(), counts))
        out = []
        for k in range(K):
            t = torch.zeros(counts[k], 2, dtype=torch.int32)
            if counts[k]:
                check(lib.scn_submanifold_rulebook_copy(self._h, i64x3(spatial_size), fs, k,
                                                        c_void_p(t.data_ptr()), stream()))
            out.append(t)
        return out

    def getRuleBook(self, in_size, out_size, filter_size, filter_stride):
        fs, st = i64x3(filter_size), i64x3(filter_stride)
        K = int(fs[0] * fs[1] * fs[2])
        counts = (c_int64 * K)()
        n_out = c_int64()
        check(lib.scn_conv_rulebook_prepare(self._h, i64x3(in_size), i64x3(out_size), fs, st,
                                            stream(), byref(n_out), counts))
        out = []
        for k in range(K):
            t = torch.zeros(counts[k], 2, dtype=torch.int32)
            if counts[k]:
                check(lib.scn_conv_rulebook_copy(self._h, i64x3(in_size), fs, st, k,
                                                 c_void_p(t.data_ptr()), stream()))
            out.append(t)
        return out

    def getSparseToDenseRuleBook(self, spatial_size):
        n = self.getNActive(spatial_size)
        rules = torch.zeros(max(n, 0), 2, dtype=torch.int32)
        sample = torch.zeros(max(n, 0), dtype=torch.int32)
        if n > 0:
            check(lib.scn_sparse_to_dense_rules_copy(self._h, i64x3(spatial_size),
                                                     c_void_p(rules.data_ptr()),
                                                     c_void_p(sample.data_ptr()), stream()))
        return rules, sample

    def ruleBookStats(self, kind, in_size, filter_size, filter_stride=(1, 1, 1)):
        st = (c_int64 * 3)()
        check(lib.scn_rulebook_stats(self._h, kind, i64x3(in_size), i64x3(filter_size),
                                     i64x3(filter_stride), st))
        return {"pairs": st[0], "tile_rows": st[1], "tiles": st[2]}

    # ---- reference methods that are off the sparse3d path ---------------------------------
    def _off_path(self, *a, **k):
        raise NotImplementedError("Metadata_3: this method is not on the sparse3d backbone path "
                                  "(SURVEY.md section 8) and is not implemented in the B200 build")

    setInputSpatialSize = batchAddSample = setInputSpatialLocation = setInputSpatialLocations = _off_path
    createMetadataForDenseToSparse = sparsifyMetadata = appendMetadata = sparsifyCompare = _off_path
    addSampleFromThresholdedTensor = generateRuleBooks3s2 = generateRuleBooks2s2 = _off_path
    compareSparseHelper = copyFeaturesHelper = _off_path


def _dims(*sizes):
    return [i64x3(s) for s in sizes]


# ---- InputLayer / OutputLayer (pybind.cpp:154-170) ---------------------------------------
def InputLayer_prepare(m, spatial_size, coords, batch_size, mode):
    """the integer half of InputLayer_updateOutput (Metadata::inputLayer, Metadata.cpp:406-417): hash
    grid, first-occurrence numbering, point lists.  Returns nActive.  B200 extension: it can run ahead
    of time, on another stream / host thread (modules.InputPrefetcher)."""
    assert coords.dim() == 2 and coords.size(1) in (3, 4), coords.shape
    coords = coords.long().contiguous()
    n_active = c_int64()
    check(lib.scn_input_layer_prepare(m._h, i64x3(spatial_size), ptr(coords), coords.size(0),
                                      coords.size(1), 1 if coords.is_cuda else 0, int(batch_size),
                                      int(mode), stream(), byref(n_active)))
    m._n_points = coords.size(0)
    return n_active.value


def InputLayer_prepare_plan(m, spatial_size, coords, batch_size, mode, plan):
    """InputLayer_prepare + every rulebook of `plan` (int64 tensor [n_ops, 13], see scn_build_plan) in one
    foreign call"""
    coords = coords.long().contiguous()
    plan = plan.contiguous()
    n_active = c_int64()
    check(lib.scn_build_plan(m._h, i64x3(spatial_size), ptr(coords), coords.size(0), coords.size(1),
                             1 if coords.is_cuda else 0, int(batch_size), int(mode), c_void_p(plan.data_ptr()),
                             plan.size(0), stream(), byref(n_active)))
    m._n_points = coords.size(0)
    return n_active.value


def InputLayer_updateOutput(m, spatial_size, coords, input_features, output_features, batch_size,
                            mode, prepared_n_active=None):
    input_features = require_cuda_f32(input_features, "InputLayer features")
    if prepared_n_active is None:
        n_active = InputLayer_prepare(m, spatial_size, coords, batch_size, mode)
    else:
        n_active = prepared_n_active
    planes = input_features.size(1)
    assert input_features.size(0) == m._n_points, (input_features.shape, m._n_points)
    output_features.resize_(n_active, planes)
    check(lib.scn_input_layer_forward(m._h, ptr(input_features), ptr(output_features), planes,
                                      stream()))


def InputLayer_updateGradInput(m, d_input_features, d_output_features):
    d_output_features = require_cuda_f32(d_output_features, "InputLayer grad_output")
    hd = (c_int64 * 4)()
    check(lib.scn_input_rulebook_header(m._h, hd, stream()))
    planes = d_output_features.size(1)
    d_input_features.resize_(hd[2], planes)
    check(lib.scn_input_layer_backward(m._h, ptr(d_input_features), ptr(d_output_features), planes,
                                       stream()))


def OutputLayer_updateOutput(m, input_features, output_features):
    input_features = require_cuda_f32(input_features, "OutputLayer features")
    hd = (c_int64 * 4)()
    check(lib.scn_input_rulebook_header(m._h, hd, stream()))
    planes = input_features.size(1)
    output_features.resize_(hd[2], planes)
    check(lib.scn_output_layer_forward(m._h, ptr(input_features), ptr(output_features), planes,
                                       stream()))


def OutputLayer_updateGradInput(m, d_input_features, d_output_features):
    d_output_features = require_cuda_f32(d_output_features, "OutputLayer grad_output")
    hd = (c_int64 * 4)()
    check(lib.scn_input_rulebook_header(m._h, hd, stream()))
    planes = d_output_features.size(1)
    d_input_features.resize_(hd[3], planes)
    check(lib.scn_output_layer_backward(m._h, ptr(d_input_features), ptr(d_output_features), planes,
                                        stream()))


# ---- convolutions (pybind.cpp:54-65,78-89,134-143) --------------------------------------
def _planes(weight):
    # weight [K, groups, nIn/groups, nOut/groups]; groups == 1 on this path
    if weight.dim() != 4 or weight.size(1) != 1:
        raise NotImplementedError("grouped convolutions (groups != 1) are not on the sparse3d path")
    return weight.size(2), weight.size(3)


def SubmanifoldConvolution_updateOutput(spatial_size, filter_size, m, input_features,
                                        output_features, weight, bias):
    x = require_cuda_f32(input_features, "input_features")
    w = require_cuda_f32(weight, "weight")
    cin, cout = _planes(w)
    ss, fs = _dims(spatial_size, filter_size)
    # a submanifold convolution keeps the active set: one output row per input row (the rulebook is
    # built, or found in the Metadata's cache, inside the forward call)
    output_features.resize_(x.size(0), cout)
    macs = c_double()
    check(lib.scn_submanifold_conv_forward(m._h, ss, fs, ptr(x), ptr(output_features), ptr(w),
                                           ptr(bias), cin, cout, _lib.precision(), stream(),
                                           byref(macs), _lib.weight_tag(weight)))
    return macs.value


def SubmanifoldConvolution_backward(spatial_size, filter_size, m, input_features, d_input_features,
                                    d_output_features, weight, d_weight, d_bias):
    x = require_cuda_f32(input_features, "input_features")
    dy = require_cuda_f32(d_output_features, "d_output_features")
    w = require_cuda_f32(weight, "weight")
    cin, cout = _planes(w)
    ss, fs = _dims(spatial_size, filter_size)
    d_input_features.resize_(x.size(0), cin)
    check(lib.scn_submanifold_conv_backward(m._h, ss, fs, ptr(x), ptr(d_input_features), ptr(dy),
                                            ptr(w), ptr(d_weight), ptr(d_bias), cin, cout,
                                            _lib.precision(), stream(), _lib.weight_tag(weight)))


def _strided(fwd, prepare_in, prepare_out, in_size, out_size, filter_size, filter_stride, m, x,
             output_features, weight, bias, n_out_of):
    x = require_cuda_f32(x, "input_features")
    w = require_cuda_f32(weight, "weight")
    cin, cout = _planes(w)
    i_s, o_s, fs, st = _dims(in_size, out_size, filter_size, filter_stride)
    n_new = c_int64()
    check(lib.scn_conv_rulebook_prepare(m._h, i64x3(prepare_in), i64x3(prepare_out), fs, st,
                                        stream(), byref(n_new), None))
    # convolution: rows of the grid the rulebook just created / found; deconvolution: the fine grid's
    output_features.resize_(n_new.value if n_out_of is prepare_out else m.getNActive(n_out_of), cout)
    macs = c_double()
    check(fwd(m._h, i_s, o_s, fs, st, ptr(x), ptr(output_features), ptr(w), ptr(bias), cin, cout,
              _lib.precision(), stream(), byref(macs), _lib.weight_tag(weight)))
    return macs.value


def Convolution_updateOutput(in_size, out_size, filter_size, filter_stride, m, input_features,
                             output_features, weight, bias):
    return _strided(lib.scn_conv_forward, in_size, out_size, in_size, out_size, filter_size,
                    filter_stride, m, input_features, output_features, weight, bias, out_size)


def Deconvolution_updateOutput(in_size, out_size, filter_size, filter_stride, m, input_features,
                               output_features, weight, bias):
    # rulebook of the down-convolution fine(out_size) -> coarse(in_size), CPU/Deconvolution.cpp:15-16
    return _strided(lib.scn_deconv_forward, out_size, in_size, in_size, out_size, filter_size,
                    filter_stride, m, input_features, output_features, weight, bias, out_size)


def _strided_backward(bwd, in_size, out_size, filter_size, filter_stride, m, input_features,
                      d_input_features, d_output_features, weight, d_weight, d_bias):
    x = require_cuda_f32(input_features, "input_features")
    dy = require_cuda_f32(d_output_features, "d_output_features")
    w = require_cuda_f32(weight, "weight")
    cin, cout = _planes(w)
    i_s, o_s, fs, st = _dims(in_size, out_size, filter_size, filter_stride)
    d_input_features.resize_(x.size(0), cin)
    check(bwd(m._h, i_s, o_s, fs, st, ptr(x), ptr(d_input_features), ptr(dy), ptr(w), ptr(d_weight),
              ptr(d_bias), cin, cout, _lib.precision(), stream(), _lib.weight_tag(weight)))


def Convolution_backward(in_size, out_size, filter_size, filter_stride, m, input_features,
                         d_input_features, d_output_features, weight, d_weight, d_bias):
    _strided_backward(lib.scn_conv_backward, in_size, out_size, filter_size, filter_stride, m,
                      input_features, d_input_features, d_output_features, weight, d_weight, d_bias)


def Deconvolution_backward(in_size, out_size, filter_size, filter_stride, m, input_features,
                           d_input_features, d_output_features, weight, d_weight, d_bias):
    _strided_backward(lib.scn_deconv_backward, in_size, out_size, filter_size, filter_stride, m,
                      input_features, d_input_features, d_output_features, weight, d_weight, d_bias)


# ---- NetworkInNetwork (sparseconvnet.h:50-60) --------------------------------------------
def NetworkInNetwork_updateOutput(input_features, output_features, weight, bias):
    x = require_cuda_f32(input_features, "input_features")
    w = require_cuda_f32(weight, "weight")
    output_features.resize_(x.size(0), w.size(1))
    macs = c_double()
    check(lib.scn_nin_forward(ptr(x), ptr(output_features), ptr(w), ptr(bias), x.size(0), w.size(0),
                              w.size(1), _lib.precision(), stream(), byref(macs)))
    return macs.value


def NetworkInNetwork_updateGradInput(d_input_features, d_output_features, weight):
    dy = require_cuda_f32(d_output_features, "d_output_features")
    w = require_cuda_f32(weight, "weight")
    d_input_features.resize_(dy.size(0), w.size(0))
    check(lib.scn_nin_backward(None, ptr(d_input_features), ptr(dy), ptr(w), None, None, dy.size(0),
                               w.size(0), w.size(1), _lib.precision(), stream()))


def NetworkInNetwork_accGradParameters(input_features, d_output_features, d_weight, d_bias):
    x = require_cuda_f32(input_features, "input_features")
    dy = require_cuda_f32(d_output_features, "d_output_features")
    # weight pointer is only read for dX, which is skipped here (d_in == NULL)
    check(lib.scn_nin_backward(ptr(x), None, ptr(dy), ptr(d_weight), ptr(d_weight), ptr(d_bias),
                               x.size(0), d_weight.size(0), d_weight.size(1), _lib.precision(),
                               stream()))


# ---- BatchNormalization (sparseconvnet.h:21-32) ------------------------------------------
def BatchNormalization_updateOutput(input_features, output_features, saveMean, saveInvStd,
                                    runningMean, runningVar, weight, bias, eps, momentum, train,
                                    leakiness):
    x = require_cuda_f32(input_features, "input_features")
    output_features.resize_as_(x)
    if x.dim() != 2:
        return
    for t, nm in ((saveMean, "saveMean"), (saveInvStd, "saveInvStd"), (runningMean, "runningMean"),
                  (runningVar, "runningVar")):
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("BatchNormalization: %s must be a contiguous CUDA tensor" % nm)
    check(lib.scn_batchnorm_forward(ptr(x), ptr(output_features), ptr(saveMean), ptr(saveInvStd),
                                    ptr(runningMean), ptr(runningVar), ptr(weight), ptr(bias),
                                    float(eps), float(momentum), 2 if train == 2 else (1 if train else 0),
                                    float(leakiness), x.size(0), x.size(1), stream()))


def BatchNormalization_backward(input_features, d_input_features, output_features,
                                d_output_features, saveMean, saveInvStd, runningMean, runningVar,
                                weight, bias, d_weight, d_bias, leakiness):
    x = require_cuda_f32(input_features, "input_features")
    y = require_cuda_f32(output_features, "output_features")
    dy = require_cuda_f32(d_output_features, "d_output_features")
    d_input_features.resize_as_(x)
    if x.dim() != 2:
        return
    check(lib.scn_batchnorm_backward(ptr(x), ptr(d_input_features), ptr(y), ptr(dy), ptr(saveMean),
                                     ptr(saveInvStd), ptr(weight), ptr(d_weight), ptr(d_bias),
                                     float(leakiness), x.size(0), x.size(1), stream()))


# ---- SparseToDense (pybind.cpp:124-133) --------------------------------------------------
def SparseToDense_updateOutput(spatial_size, m, input_features, output_features, n_planes):
    x = require_cuda_f32(input_features, "input_features")
    ss = i64x3(spatial_size)
    batch = m.getBatchSize()
    output_features.resize_(batch, int(n_planes), int(ss[0]), int(ss[1]), int(ss[2]))
    check(lib.scn_sparse_to_dense_forward(m._h, ss, ptr(x) if x.dim() == 2 else None,
                                          ptr(output_features), int(n_planes), batch, stream()))


def SparseToDense_updateGradInput(spatial_size, m, input_features, d_input_features,
                                  d_output_features):
    dy = require_cuda_f32(d_output_features, "d_output_features")
    d_input_features.resize_as_(input_features)
    d_input_features.zero_()
    if input_features.dim() != 2:
        return
    check(lib.scn_sparse_to_dense_backward(m._h, i64x3(spatial_size), ptr(d_input_features), ptr(dy),
                                           input_features.size(1), dy.size(0), stream()))


# B200 extension: the occupied extent of a grid and SparseToDense into the cropped volume (tools_3d_2d.py:7-48)
def grid_extent(m, spatial_size):
    """[max x + 1, max y + 1, max z + 1, max batch index + 1] of the active sites at `spatial_size`"""
    ext = (c_int64 * 4)()
    check(lib.scn_grid_extent(m._h, i64x3(spatial_size), ext, stream()))
    return [int(v) for v in ext]


def SparseToDense_cropped_updateOutput(spatial_size, extent, m, input_features, output_features, n_planes):
    x = require_cuda_f32(input_features, "input_features")
    batch = m.getBatchSize()
    output_features.resize_(batch, int(n_planes), int(extent[0]), int(extent[1]), int(extent[2]))
    check(lib.scn_sparse_to_dense_cropped_forward(m._h, i64x3(spatial_size), i64x3(extent), ptr(x) if x.dim() == 2 else None,
                                                  ptr(output_features), int(n_planes), batch, stream()))


def SparseToDense_cropped_updateGradInput(spatial_size, extent, m, input_features, d_input_features, d_output_features):
    dy = require_cuda_f32(d_output_features, "d_output_features")
    d_input_features.resize_as_(input_features)
    d_input_features.zero_()
    if input_features.dim() != 2:
        return
    check(lib.scn_sparse_to_dense_cropped_backward(m._h, i64x3(spatial_size), i64x3(extent), ptr(d_input_features), ptr(dy),
                                                   input_features.size(1), dy.size(0), stream()))


# ---- instrumentation -----------------------------------------------------------------------
def launch_count():
    """kernels launched by libscn_b200 since load (bench.py `gpu_launches`)"""
    return int(lib.scn_launch_count())


PROF_CLASSES = ("conv_gemm", "weight_grad", "batchnorm", "rulebook", "io")


def prof_enable(on):
    check(lib.scn_prof_enable(1 if on else 0))


def prof_read():
    """{class: dict(regions, ms, bytes, flops)} accumulated since the last read (synchronises)"""
    out = {}
    for i, name in enumerate(PROF_CLASSES):
        v = (c_double * 4)()
        check(lib.scn_prof_read(i, v))
        out[name] = {"regions": int(v[0]), "ms": v[1], "bytes": v[2], "flops": v[3]}
    return out


# ---- extension entry points off the sparse3d path --------------------------------------------
_OP_SUFFIXES = ("_updateOutput", "_updateGradInput", "_backward", "_accGradParameters")


def __getattr__(name):
    """The reference's extension also exports the pooling / full / randomized-stride / permutohedral /
    BL-layer entry points (pybind.cpp:34-235).  None of them is on the sparse3d path; a
    layer file that reaches for one gets a loud NotImplementedError at call time instead of an AttributeError
    at some unrelated line."""
    if name.endswith(_OP_SUFFIXES) and not name.startswith("_"):
        def _off_path(*a, **k):
            raise NotImplementedError("sparseconvnet.SCN.%s is not on the sparse3d backbone path (SURVEY.md "
                                      "section 8) and is not implemented in the B200 build" % name)
        _off_path.__name__ = name
        _off_path.off_path = True
        return _off_path
    raise AttributeError("module 'sparseconvnet.SCN' has no attribute %r" % name)
