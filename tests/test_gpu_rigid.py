"""Dynamic rigid bodies on the CUDA path against the oracle (wrench atomics, renew_rigid_particle_state);
reference: base_solver.py:174-187, 272-278, 615-629, DFSPH.py:190-202, 270-283."""
import numpy as np
import pytest

from helpers import by_uid, make_sim, oracle_library
from test_rigid_stepper import cube_scene

pytestmark = pytest.mark.gpu


def test_dynamic_cube_matches_oracle(tmp_path):
    sc = cube_scene(tmp_path, fluid=True, cube_y=0.37, cube_v=(0.0, -2.0, 0.0), density=500.0)
    (cg, sg), (co, so) = make_sim(sc), make_sim(sc, oracle_library())
    for _ in range(30):
        sg.step(), so.step()
    bg, bo = sg.rigid_solver.bodies[1], so.rigid_solver.bodies[1]
    assert np.allclose(bg.x, bo.x, atol=2e-4) and np.allclose(bg.v, bo.v, atol=2e-2)
    xg, xo = by_uid(cg, cg.particle_positions), by_uid(co, co.particle_positions)
    assert np.abs(xg - xo).max() / np.abs(xo).max() < 1e-3
