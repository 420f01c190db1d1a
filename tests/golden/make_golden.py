"""Generates tests/golden/*.npz with the CPU oracle (oracle/sph_oracle.cpp).

The reference cannot be imported here (Taichi / pybullet / trimesh are not installable), so these are
NOT reference outputs: they freeze the oracle's own results on small seeded scenes so that neither the
oracle nor the CUDA path can drift unnoticed.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import by_uid, make_sim, oracle_library, scene  # noqa: E402

CASES = {
    # name: (method, dt, steps, scene kwargs)
    "dfsph_dam_break_2k": ("dfsph", 1e-3, 40, {}),
    "wcsph_dam_break_2k": ("wcsph", 4e-4, 60, {}),
    "pcisph_dam_break_2k": ("pcisph", 8e-4, 40, {}),
    "dfsph_implicit_viscosity_2k": ("dfsph", 1e-3, 20, dict(viscosity_method="implicit", viscosity=50.0, viscosity_b=50.0)),
}


def case_scene(method, dt, kw):
    return scene(method, domain_end=(0.6, 0.8, 0.6), block_start=(0.1, 0.1, 0.1), block_end=(0.3, 0.5, 0.3),
                 velocity=(0.0, -1.0, 0.0), dt=dt, **kw)


def run_case(name, lib):
    method, dt, steps, kw = CASES[name]
    c, s = make_sim(case_scene(method, dt, kw), lib)
    st = s.step(steps)
    mat = by_uid(c, c.particle_materials)
    return dict(positions=by_uid(c, c.particle_positions)[mat == 1], velocities=by_uid(c, c.particle_velocities)[mat == 1],
                densities=by_uid(c, c.particle_densities)[mat == 1],
                iterations=np.array([st.total_dfsph_iterations, st.total_dfsph_iterations_v, st.total_pcisph_iterations,
                                     st.total_cg_iterations], dtype=np.int64))


if __name__ == "__main__":
    lib = oracle_library()
    for name in CASES:
        out = run_case(name, lib)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()}, out["iterations"])
