from .base_container import BaseContainer
from .dfsph_container import DFSPHContainer
from .pcisph_container import PCISPHContainer
from .wcsph_container import WCSPHContainer

__all__ = ["BaseContainer", "DFSPHContainer", "PCISPHContainer", "WCSPHContainer"]
