"""Built-in rigid-body stepper behind the PyBulletSolver boundary, on the CPU oracle
engine: free fall, wall contact, and two-way coupling with DFSPH (signs and momentum exchange).
PyBullet itself is not installable, so trajectories cannot be pinned against it."""
import numpy as np

from helpers import by_uid, make_sim, oracle_library, scene
from test_mesh import write_box_obj


def cube_scene(tmp_path, method="dfsph", fluid=True, cube_y=0.5, cube_v=(0.0, 0.0, 0.0), density=500.0, dt=1e-3):
    obj = tmp_path / "cube.obj"
    write_box_obj(obj, [-0.05, -0.05, -0.05], [0.05, 0.05, 0.05])
    sc = scene(method, dt=dt, domain_end=(0.6, 1.0, 0.6), block_start=(0.1, 0.1, 0.1), block_end=(0.5, 0.3, 0.5))
    if not fluid:
        sc["FluidBlocks"] = []
    sc["RigidBodies"] = [dict(objectId=1, geometryFile=str(obj), translation=[0.3, cube_y, 0.3], rotationAxis=[0, 1, 0],
                              rotationAngle=0, scale=[1, 1, 1], velocity=list(cube_v), density=density, color=[255, 0, 0],
                              isDynamic=True, entryTime=-1.0)]
    return sc


def test_free_fall_and_floor_contact(tmp_path):
    c, s = make_sim(cube_scene(tmp_path, fluid=False, cube_y=0.5), oracle_library())
    rs = s.rigid_solver
    assert not rs.is_noop and 1 in rs.bodies
    n = 50
    for _ in range(n):
        s.step()
    t = n * 1e-3
    b = rs.bodies[1]
    # semi-implicit Euler; dt reaches the stepper as the f32 the solver stores
    assert abs(b.v[1] - (-9.81 * t)) < 1e-6 and abs(b.x[1] - (0.5 - 0.5 * 9.81 * t * (t + 1e-3))) < 1e-6
    obj = by_uid(c, c.particle_object_ids)
    x = by_uid(c, c.particle_positions)[obj == 1]
    assert abs(x[:, 1].mean() - b.x[1]) < 1e-3            # particles follow the body state
    for _ in range(400):                                   # until it sits on the inset floor wall
        s.step()
    b = rs.bodies[1]
    floor = c.padding + c.particle_diameter + c.domain_box_thickness
    assert abs((b.x[1] - 0.05) - floor) < 2e-3 and abs(b.v[1]) < 0.05
    assert np.allclose(b.R @ b.R.T, np.eye(3), atol=1e-9)


def test_two_way_coupling_with_dfsph(tmp_path):
    """A light cube driven into the fluid is decelerated by it, and the fluid gains momentum."""
    sc = cube_scene(tmp_path, fluid=True, cube_y=0.37, cube_v=(0.0, -2.0, 0.0), density=500.0)
    c, s = make_sim(sc, oracle_library())
    n = 40
    for _ in range(n):
        s.step()
    b = s.rigid_solver.bodies[1]
    free_fall_v = -2.0 - 9.81 * n * 1e-3
    assert b.v[1] > free_fall_v + 0.2                      # the fluid pushed back
    mat = by_uid(c, c.particle_materials)
    obj = by_uid(c, c.particle_object_ids)
    v = by_uid(c, c.particle_velocities)
    assert np.isfinite(v).all()
    under = (mat == 1) & (np.abs(by_uid(c, c.particle_positions)[:, 0] - 0.3) < 0.05) & \
            (np.abs(by_uid(c, c.particle_positions)[:, 2] - 0.3) < 0.05)
    assert v[under, 1].mean() < -9.81 * n * 1e-3 - 0.05    # fluid under the cube moves down faster than free fall
    assert (obj == 1).sum() == c.rigid_bodies[0]["particleNum"]


def test_initial_orientation_is_euler_xyz_like_upstream():
    """Upstream feeds rotationAxis * angle to p.getQuaternionFromEuler (bullet_solver.py:102-106): roll, pitch, yaw."""
    from sph_project_b200.rigid_solver.bullet_solver import _euler_xyz, _skew_exp
    for axis in np.eye(3):                       # coordinate axes: Euler and axis-angle agree
        assert np.allclose(_euler_xyz(axis * 0.7), _skew_exp(axis, 0.7), atol=1e-12)
    rpy = np.array([0.3, -0.5, 0.9])
    R = _euler_xyz(rpy)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
    Rx, Ry, Rz = (_skew_exp(a * t, 1.0) for a, t in zip(np.eye(3), rpy))
    assert np.allclose(R, Rz @ Ry @ Rx, atol=1e-12)
    oblique = np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0) * 0.8
    assert not np.allclose(_euler_xyz(oblique), _skew_exp(oblique / 0.8, 0.8), atol=1e-3)   # they differ off the axes
