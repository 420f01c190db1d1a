"""Drop-in import paths of the reference: `from SPH.utils import SimConfig`,
`from SPH.containers import DFSPHContainer`, `from SPH.fluid_solvers import DFSPHSolver`
(run_simulation.py:5-7).  Everything is implemented in `sph_project_b200`."""
import importlib
import sys

for _name in ("utils", "containers", "fluid_solvers", "rigid_solver"):
    _mod = importlib.import_module(f"sph_project_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
del _name, _mod
