"""sparseconvnet for B200: the SparseConvNet module API of xuyongzhi/Automatic-As-built-Reconstruction
(SparseConvNet/sparseconvnet/__init__.py:13-42) restricted to what the sparse3d backbone reaches,
running on hand-written sm_100a CUDA behind a C ABI (include/scn_b200.h).  No CPU fallback."""
forward_pass_multiplyAdd_count = 0
forward_pass_hidden_states = 0

from ._lib import set_conv_precision, get_conv_precision, LIB_PATH  # noqa: E402
from . import SCN  # noqa: E402
from .tensor import SparseConvNetTensor  # noqa: E402
from .modules import (  # noqa: E402
    AddTable, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, CheckpointedSequential,
    ConcatTable, Convolution, Deconvolution, Identity, InputLayer, InputLayerInput, InputPrefetcher,
    JoinTable, PreparedInput,
    Metadata, NetworkInNetwork, OutputLayer, Sequential, SparseToDense, SubmanifoldConvolution,
    ValidConvolution, add_feature_planes, concatenate_feature_planes, optionalTensor,
    optionalTensorReturn, toLongTensor)
from . import (batchNormalization, convolution, deconvolution, identity, ioLayers, metadata,  # noqa: E402
               networkInNetwork, sequential, sparseConvNetTensor, sparseToDense,
               submanifoldConvolution, tables, utils)
from .graph import GraphFunction, LayerGraph  # noqa: E402
from .fpn_net import FPN_Net  # noqa: E402
from . import tools_3d_2d  # noqa: E402
from .roi_align_rotated_3d import ROIAlignRotated3D, roi_align_rotated_3d  # noqa: E402
from .rpn import RPNHead, grid_anchors  # noqa: E402
from .nms import boxes_iou_3d, rotate_iou_gpu_eval, rotate_nms, rotate_nms_3d  # noqa: E402
from .voxelize import VoxelLoader, quantize_points, voxelize_batch  # noqa: E402
from .data_parallel import GradBucket, broadcast_parameters, shard_indices  # noqa: E402


def _off_path(name):
    class _OffPath(object):
        def __init__(self, *a, **k):
            raise NotImplementedError(
                "sparseconvnet.%s is not reachable from the sparse3d backbone (FPN_Net) and is out of "
                "scope of the B200 build (SURVEY.md section 2, row 13)" % name)
    _OffPath.__name__ = name
    return _OffPath


# exported by the reference __init__ but never instantiated on the sparse3d path
for _n in ("Tanh", "Sigmoid", "ReLU", "LeakyReLU", "ELU", "SELU", "BatchNormELU", "AveragePooling",
           "MeanOnlyBNLeakyReLU", "ClassificationTrainValidate", "DenseToSparse", "Dropout",
           "BatchwiseDropout", "FullConvolution", "TransposeConvolution", "InputBatch", "BLInputLayer",
           "BLOutputLayer", "MaxPooling", "PermutohedralSubmanifoldConvolution",
           "RandomizedStrideConvolution", "RandomizedStrideMaxPooling", "Sparsify", "SparsifyFCS",
           "UnPooling", "ShapeContext", "MultiscaleShapeContext", "AddCoords"):
    globals()[_n] = _off_path(_n)
del _n
