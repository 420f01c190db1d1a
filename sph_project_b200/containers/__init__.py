"""Particle containers behind the reference's `SPH.containers` names: host-side views of the device
state owned by the C ABI handle (include/sph_b200.h).  IISPH / PBF containers are not provided (their
solvers are broken upstream)."""
from . import base_container, dfsph_container, pcisph_container, wcsph_container

BaseContainer = base_container.BaseContainer
WCSPHContainer = wcsph_container.WCSPHContainer
PCISPHContainer = pcisph_container.PCISPHContainer
DFSPHContainer = dfsph_container.DFSPHContainer

__all__ = ["BaseContainer", "DFSPHContainer", "PCISPHContainer", "WCSPHContainer"]
