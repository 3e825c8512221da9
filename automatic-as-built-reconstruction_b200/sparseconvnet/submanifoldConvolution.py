"""compat import path: sparseconvnet.submanifoldConvolution (reference file of the same name)."""
from .modules import SubmanifoldConvolution, ValidConvolution, SubmanifoldConvolutionFunction  # noqa: F401
