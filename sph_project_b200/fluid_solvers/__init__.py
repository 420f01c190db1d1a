from .base_solver import BaseSolver
from .DFSPH import DFSPHSolver
from .PCISPH import PCISPHSolver
from .WCSPH import WCSPHSolver

__all__ = ["BaseSolver", "DFSPHSolver", "PCISPHSolver", "WCSPHSolver"]
