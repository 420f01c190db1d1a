"""Drop-in surface: every public name the reference's container / solver / rigid-solver / config objects carry
(tests/golden/ref_api_surface.json, extracted from the reference's own classes by
tests/golden/make_ref_api_surface.py) exists on this repository's classes with the same kind
(host-callable method or kernel / field with to_numpy / plain value), except the plumbing listed below.
`@ti.func`s are device-side only upstream (Taichi refuses to call them from Python scope): their equivalents are
compiled into the library; the handful that is useful on the host (kernel_W, kernel_gradient, pos_to_index, ...) is
offered as numpy helpers and checked against the oracle here."""
import json
import os

import pytest

from helpers import ROOT, make_sim, oracle_library, scene

# Taichi / PyBullet plumbing that has no meaning behind the C ABI (sort buffers, the scan executor, the temp
# counters, Bullet's id maps and helpers); everything else must be there.
NOT_PROVIDED = {
    "container": {"grid_num_particles_temp", "grid_ids_buffer", "grid_ids_new", "is_dynamic_buffer",
                  "particle_colors_buffer", "particle_densities_buffer", "particle_masses_buffer", "particle_materials_buffer",
                  "particle_object_ids_buffer", "particle_positions_buffer", "particle_rest_volumes_buffer",
                  "particle_velocities_buffer", "rigid_particle_original_positions_buffer"},
    "solver": {"add_viscosity_force_to_rigid"},       # dead code upstream: defined (base_solver.py:476), never called
    "rigid_solver": {"container_idx_to_bullet_idx", "bullet_idx_to_container_idx", "create_wall", "init_rigid_block"},
    "config": set(),
}

SURFACE = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_api_surface.json")))


def kind(obj, name):
    v = getattr(obj, name)
    if callable(v):
        return "callable"
    if hasattr(v, "to_numpy"):
        return "field"
    return "value"


@pytest.mark.parametrize("method", ["wcsph", "pcisph", "dfsph"])
def test_reference_names_are_provided(method):
    c, s = make_sim(scene(method, domain_end=(0.6, 0.6, 0.6), block_start=(0.2, 0.2, 0.2), block_end=(0.3, 0.3, 0.3)),
                    oracle_library(), prepare=False, GGUI=True)
    ours = {"container": c, "solver": s, "rigid_solver": s.rigid_solver, "config": c.cfg}
    missing, wrong_kind = [], []
    for part, names in SURFACE[method].items():
        for name, ref_kind in names.items():
            if name in NOT_PROVIDED[part] or name.startswith("_") or ref_kind == "ti.func":
                continue
            want = "callable" if ref_kind in ("method", "ti.kernel") else ref_kind
            if not hasattr(ours[part], name):
                missing.append(f"{part}.{name} ({ref_kind})")
            elif want in ("callable", "field") and kind(ours[part], name) != want:
                wrong_kind.append(f"{part}.{name}: {ref_kind} upstream, {kind(ours[part], name)} here")
    assert not missing, missing
    assert not wrong_kind, wrong_kind


def test_host_helpers_for_device_functions():
    """kernel_W / kernel_gradient / cell helpers against closed forms and the oracle's own sort."""
    import numpy as np
    c, s = make_sim(scene("wcsph", domain_end=(0.6, 0.6, 0.6), block_start=(0.2, 0.2, 0.2), block_end=(0.3, 0.3, 0.3)),
                    oracle_library())
    h = c.dh
    assert np.isclose(s.kernel_W(0.0), 8 / np.pi / h ** 3, rtol=1e-6) and s.kernel_W(h) == 0 and s.kernel_W(1.5 * h) == 0
    r = np.array([0.3 * h, 0.0, 0.0], dtype=np.float32)
    g = s.kernel_gradient(r)
    eps = 1e-4 * h
    fd = (s.kernel_W(0.3 * h + eps).astype(np.float64) - s.kernel_W(0.3 * h - eps)) / (2 * eps)
    assert np.isclose(g[0], fd, rtol=2e-2) and g[1] == 0 and np.all(s.kernel_gradient(np.zeros(3)) == 0)
    assert np.allclose(s.kernel_gradient(-r), -g)
    n = c.particle_num[None]
    x = c.particle_positions.to_numpy(n)
    assert np.array_equal(c.get_flatten_grid_index(x), c.grid_ids.to_numpy(n))          # the oracle's own cell ids
    mat, dyn = c.particle_materials.to_numpy(n), c.particle_is_dynamic.to_numpy(n)
    p = int(np.flatnonzero(mat == 2)[0])
    assert c.is_static_rigid_body(p) and not c.is_dynamic_rigid_body(p)
    com = c.compute_rigid_body_center_of_mass(0)
    fluid = (c.particle_object_ids.to_numpy(n) == 0) & (dyn != 0)
    assert np.allclose(com, x[fluid].mean(0), atol=1e-5)
    buf = np.zeros((c.particle_max_num, 3), dtype=np.float32)
    c.copy_to_numpy(buf, c.particle_positions)
    assert np.array_equal(buf[:n], x)
    c.init_grid(); c.prefix_sum_executor.run(c.grid_num_particles); c.reorder_particles()      # upstream's three-call sequence
    assert np.array_equal(c.particle_positions.to_numpy(n), x)
