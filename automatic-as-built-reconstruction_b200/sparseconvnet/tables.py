"""compat import path: sparseconvnet.tables (reference file of the same name)."""
from .modules import JoinTable, AddTable, ConcatTable  # noqa: F401
