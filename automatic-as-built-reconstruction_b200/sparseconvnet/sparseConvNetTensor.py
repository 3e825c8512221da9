"""compat import path: sparseconvnet.sparseConvNetTensor."""
from .tensor import SparseConvNetTensor  # noqa: F401
