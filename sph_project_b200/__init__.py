"""B200-native SPH inner loop behind the SPH_Project container/solver API.

    from sph_project_b200.utils import SimConfig
    from sph_project_b200.containers import DFSPHContainer
    from sph_project_b200.fluid_solvers import DFSPHSolver

The top-level `SPH` package of this repository aliases these modules under the reference's
import paths (`SPH.utils`, `SPH.containers`, `SPH.fluid_solvers`, `SPH.rigid_solver`).
"""
from . import _native
from ._native import CUDA_LIBRARY_PATH, SphError, load_cuda_library

__all__ = ["_native", "CUDA_LIBRARY_PATH", "SphError", "load_cuda_library"]
__version__ = "0.1.0"
