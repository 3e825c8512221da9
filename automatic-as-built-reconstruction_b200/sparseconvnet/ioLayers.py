"""compat import path: sparseconvnet.ioLayers (reference file of the same name)."""
from .modules import InputLayer, OutputLayer, InputLayerInput, InputLayerFunction, OutputLayerFunction  # noqa: F401
