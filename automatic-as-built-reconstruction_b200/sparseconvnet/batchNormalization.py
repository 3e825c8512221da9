"""compat import path: sparseconvnet.batchNormalization (reference file of the same name)."""
from .modules import BatchNormalization, BatchNormReLU, BatchNormLeakyReLU, BatchNormalizationFunction  # noqa: F401
