"""compat import path: sparseconvnet.identity (reference file of the same name)."""
from .modules import Identity  # noqa: F401
