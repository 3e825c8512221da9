"""compat import path: sparseconvnet.networkInNetwork (reference file of the same name)."""
from .modules import NetworkInNetwork, NetworkInNetworkFunction  # noqa: F401
