"""compat import path: sparseconvnet.metadata (reference file of the same name)."""
from .modules import Metadata  # noqa: F401
