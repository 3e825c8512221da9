"""compat import path: sparseconvnet.sequential (reference file of the same name)."""
from .modules import Sequential, CheckpointedSequential  # noqa: F401
