"""compat import path: sparseconvnet.sparseToDense (reference file of the same name)."""
from .modules import SparseToDense, SparseToDenseFunction  # noqa: F401
