"""Shared test plumbing: synthetic scenes, the oracle library, uid-ordered field access."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libsph_oracle.so")

_oracle = None


def oracle_library():
    """The CPU oracle (test infrastructure), built on demand with its own Makefile."""
    global _oracle
    if _oracle is None:
        src = os.path.join(ORACLE_DIR, "sph_oracle.cpp")
        if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
        from sph_project_b200 import _native
        _oracle = _native.bind(ctypes.CDLL(ORACLE_LIB))
    return _oracle


def scene(method="wcsph", domain_end=(1.0, 1.0, 1.0), block_start=(0.1, 0.1, 0.1), block_end=(0.495, 0.495, 0.495),
          velocity=(0.0, 0.0, 0.0), dt=4e-4, viscosity_method="standard", viscosity=10.0, viscosity_b=5.0,
          add_domain_box=True, radius=0.01, g_upper=None, extra_blocks=(), density0=1000.0, translation=(0, 0, 0)):
    """Scene dict in the reference's JSON schema.  Defaults = config C1."""
    cfg = {
        "domainStart": [0.0, 0.0, 0.0], "domainEnd": list(domain_end), "particleRadius": radius,
        "addDomainBox": add_domain_box, "density0": density0, "gravitation": [0.0, -9.81, 0.0],
        "simulationMethod": method, "viscosityMethod": viscosity_method, "viscosity": viscosity,
        "viscosity_b": viscosity_b, "timeStepSize": dt, "exportFrame": False, "exportPly": False, "exportObj": False,
    }
    if g_upper is not None:
        cfg["gravitationUpper"] = g_upper
    blocks = [{"objectId": 0, "start": list(block_start), "end": list(block_end), "translation": list(translation),
               "scale": [1, 1, 1], "velocity": list(velocity), "density": 1000.0, "color": [50, 100, 200],
               "entryTime": -1.0}]
    blocks += list(extra_blocks)
    return {"Configuration": cfg, "FluidBlocks": blocks}


def make_sim(scene_dict, lib=None, prepare=True, device=0, slab=None, GGUI=False):
    """(container, solver) for a scene; lib=None -> the CUDA product, else a bound library."""
    import copy
    from sph_project_b200.containers import DFSPHContainer, PCISPHContainer, WCSPHContainer
    from sph_project_b200.fluid_solvers import DFSPHSolver, PCISPHSolver, WCSPHSolver
    from sph_project_b200.utils import SimConfig
    classes = {"wcsph": (WCSPHContainer, WCSPHSolver), "pcisph": (PCISPHContainer, PCISPHSolver),
               "dfsph": (DFSPHContainer, DFSPHSolver)}
    config = SimConfig(config=copy.deepcopy(scene_dict), verbose=False)
    C, S = classes[config.get_cfg("simulationMethod")]
    container = C(config, GGUI=GGUI, engine_library=lib, device=device, slab=slab)
    solver = S(container)
    if prepare:
        solver.prepare()
    return container, solver


def by_uid(container, field):
    """Field values of the live particles ordered by insertion index (undoes the sort)."""
    n = container.particle_num[None]
    uid = container.particle_uids.to_numpy(n)
    a = field.to_numpy(n)
    out = np.empty_like(a)
    out[uid] = a
    return out
