"""compat import path: sparseconvnet.utils (reference file of the same name)."""
from .modules import toLongTensor, optionalTensor, optionalTensorReturn, add_feature_planes, concatenate_feature_planes  # noqa: F401
