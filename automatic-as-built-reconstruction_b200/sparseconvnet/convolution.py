"""compat import path: sparseconvnet.convolution (reference file of the same name)."""
from .modules import Convolution, ConvolutionFunction  # noqa: F401
