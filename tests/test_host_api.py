"""Host-side mirror of the reference's container / solver surface (fields, scalars, errors), run on
the CPU oracle engine."""
import numpy as np
import pytest

from helpers import by_uid, make_sim, oracle_library, scene
from sph_project_b200 import _native


@pytest.fixture(scope="module")
def sim():
    return make_sim(scene("dfsph", dt=1e-3, domain_end=(0.6, 0.6, 0.6), block_start=(0.1, 0.1, 0.1), block_end=(0.3, 0.3, 0.3)),
                    oracle_library())


def test_container_attributes_match_reference_names(sim):
    c, s = sim
    for name in ("dim dx dh particle_diameter particle_spacing V0 padding grid_size grid_num domain_start domain_end "
                 "domain_size particle_max_num particle_num fluid_particle_num object_num object_collection "
                 "object_id_fluid_body object_id_rigid_body present_object total_time material_fluid material_rigid "
                 "particle_object_ids particle_positions particle_velocities particle_accelerations particle_rest_volumes "
                 "particle_masses particle_densities particle_pressures particle_materials particle_colors particle_is_dynamic "
                 "rigid_particle_original_positions grid_ids object_materials rigid_body_is_dynamic rigid_body_masses "
                 "rigid_body_centers_of_mass rigid_body_rotations rigid_body_torques rigid_body_forces rigid_body_velocities "
                 "rigid_body_angular_velocities object_visibility particle_dfsph_alphas particle_dfsph_kappa "
                 "particle_dfsph_kappa_v particle_densities_star particle_densities_derivatives").split():
        assert hasattr(c, name), name
    for name in ("prepare step _step compute_density compute_non_pressure_acceleration compute_pressure_acceleration "
                 "update_fluid_velocity update_fluid_position enforce_domain_boundary_3D renew_rigid_particle_state "
                 "compute_rigid_particle_volume compute_alpha compute_density_star compute_density_derivative compute_kappa "
                 "compute_kappa_v correct_divergence_step correct_density_error_step correct_divergence_error "
                 "correct_density_error compute_density_error compute_density_derivative_error").split():
        assert callable(getattr(s, name)), name
    assert c.material_fluid == 1 and c.material_rigid == 2 and c.dim == 3
    assert np.isclose(c.dh, 4 * c.dx) and np.isclose(c.V0, 0.8 * (2 * c.dx) ** 3)
    assert s.dt[None] == pytest.approx(1e-3) and s.density_0 == 1000.0 and s.rigid_solver.is_noop


def test_field_views(sim):
    c, s = sim
    n = c.particle_num[None]
    assert c.particle_positions.to_numpy().shape == (c.particle_max_num, 3)
    assert c.particle_densities.to_numpy(n).shape == (n,)
    v = c.particle_velocities.to_numpy(n)
    v2 = v + np.float32(0.25)
    c.particle_velocities.from_numpy(v2)
    assert np.array_equal(c.particle_velocities.to_numpy(n), v2)
    assert np.array_equal(c.particle_velocities[3], v2[3])
    c.particle_velocities[3] = [1.0, 2.0, 3.0]
    assert np.array_equal(c.particle_velocities[3], np.array([1, 2, 3], np.float32))
    c.particle_velocities.from_numpy(v)
    c.particle_pressures.fill(7.0)
    assert np.all(c.particle_pressures.to_numpy(n) == 7.0)
    # dump(): positions / velocities of one object (base_container.py:599-609)
    d = c.dump(0)
    assert d["position"].shape == (1000, 3) and d["velocity"].shape == (1000, 3)


def test_scalars_and_object_tables(sim):
    c, s = sim
    assert c.fluid_particle_num[None] == 1000 and c.object_num[None] == 2
    s.dt[None] = 5e-4
    assert s.dt[None] == pytest.approx(5e-4)
    s.dt[None] = 1e-3
    box = c.object_num[None] - 1
    assert c.object_materials[box] == c.material_rigid and c.rigid_body_is_dynamic[box] == 0
    assert np.all(c.rigid_body_forces.to_numpy() == 0)


def test_errors_are_python_exceptions():
    lib = oracle_library()
    c, s = make_sim(scene("wcsph", domain_end=(0.4, 0.4, 0.4), block_start=(0.1, 0.1, 0.1), block_end=(0.2, 0.2, 0.2)), lib)
    with pytest.raises(_native.SphError) as e:       # particle_max_num is exact: one more particle does not fit
        c.add_particles(5, 1, np.zeros((1, 3)), np.zeros((1, 3)), np.ones(1), np.zeros(1), np.ones(1, int), np.ones(1, int), np.zeros((1, 3), int))
    assert e.value.code == -2
    with pytest.raises(_native.SphError):
        c.engine.run_task(12345)
    bad = scene("wcsph")
    bad["Configuration"]["viscosityMethod"] = "nope"
    with pytest.raises(NotImplementedError):
        make_sim(bad, lib)
    bad = scene("wcsph")
    bad["RigidBlocks"] = [{}]
    with pytest.raises(NotImplementedError):
        make_sim(bad, lib)
    bad = scene("wcsph")
    bad["Configuration"]["domainStart"] = [0.0, -1.0, 0.0]
    with pytest.raises(AssertionError):
        make_sim(bad, lib)


def test_for_all_neighbors_host_callback(sim):
    c, s = sim
    off, idx = c.neighbor_lists()
    i = int(np.argmax(np.diff(off)))
    seen = c.for_all_neighbors(i, lambda p_i, p_j, ret: ret + [p_j], [])
    assert seen == list(idx[off[i]:off[i + 1]]) and i not in seen
    x = c.particle_positions.to_numpy(c.particle_num[None])
    assert all(np.linalg.norm(x[i] - x[j]) < c.dh + 1e-6 for j in seen)


def test_wrench_tables_reset_one_object_and_one_table_at_a_time():
    """The reference zeroes rigid_body_forces[i] and rigid_body_torques[i] body by body while it walks its bodies
    (bullet_solver.py:149-156); the library can only clear everything at once."""
    from sph_project_b200.fields import WrenchTable

    class FakeEngine:
        def __init__(self):
            self.f = np.arange(60, dtype=np.float32).reshape(20, 3)
            self.t = -np.arange(60, dtype=np.float32).reshape(20, 3)

        def get_rigid_wrench(self):
            return self.f.copy(), self.t.copy()

        def zero_rigid_wrench(self):
            self.f[:] = 0
            self.t[:] = 0

    eng = FakeEngine()
    forces = WrenchTable(eng, 0)
    torques = WrenchTable(eng, 1, forces.state)
    f0, t0 = forces.to_numpy().copy(), torques.to_numpy().copy()
    forces[3] = np.zeros(3)
    assert np.all(forces[3] == 0) and np.array_equal(forces[4], f0[4]) and np.array_equal(torques[3], t0[3])
    torques[3] = np.zeros(3)
    assert np.all(torques[3] == 0) and np.array_equal(torques[5], t0[5])
    eng.f[4] += 1.0                                   # the kernels keep accumulating on the device
    assert np.array_equal(forces[4], f0[4] + 1.0)
    forces.fill(0.0)
    assert not forces.to_numpy().any() and np.array_equal(torques[5], t0[5])
    torques.fill(0.0)
    assert not torques.to_numpy().any() and forces.state.remainder is None
    with pytest.raises(NotImplementedError):
        forces[1] = np.ones(3)
