"""Import-path alias: the reference tree keeps its sparse-convolution package at `SparseConvNet/sparseconvnet` and
a few of its files import it by that path (maskrcnn_benchmark/layers/roi_align_rotated_3d.py:7
`from SparseConvNet.sparseconvnet.tools_3d_2d import sparse_3d_to_dense_2d`).  `SparseConvNet.sparseconvnet` is the
B200 package itself, so those imports resolve unchanged."""
import importlib
import sys

sparseconvnet = importlib.import_module("sparseconvnet")
sys.modules[__name__ + ".sparseconvnet"] = sparseconvnet
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("sparseconvnet.") and _mod is not None:
        sys.modules[__name__ + "." + _name] = _mod
