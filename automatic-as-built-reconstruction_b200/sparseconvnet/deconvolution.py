"""compat import path: sparseconvnet.deconvolution (reference file of the same name)."""
from .modules import Deconvolution, DeconvolutionFunction  # noqa: F401
