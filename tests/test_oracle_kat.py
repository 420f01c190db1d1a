"""Known answers that pin the CPU oracle (SURVEY.md 8(c) pins 1-3; the reference has no tests)."""
import numpy as np
import pytest

from helpers import by_uid, make_sim, oracle_library, scene
from sph_project_b200._native import T


@pytest.fixture(scope="module")
def c1():
    return make_sim(scene("dfsph", dt=1e-3), oracle_library())


def test_scene_counts_c1(c1):
    c, s = c1
    # 20^3 fluid lattice + the asymmetric domain-box shell (App. B#14)
    assert c.fluid_particle_num[None] == 8000
    assert c.particle_num[None] == 8000 + 17829 == c.particle_max_num


def test_kernel_constants(c1):
    c, _ = c1
    h = c.dh
    assert np.isclose(8 / (np.pi * h ** 3), 39788.7358, rtol=1e-9)
    assert np.isclose(c.V0, 6.4e-6)
    assert np.isclose(c.V0 * 8 / (np.pi * h ** 3), 0.2546479, rtol=1e-6)


def test_lattice_density_and_alpha(c1):
    c, s = c1
    mat = by_uid(c, c.particle_materials)
    rho = by_uid(c, c.particle_densities)[mat == 1]
    # interior particle of a cubic lattice with spacing 2r: rho = 0.79998 rho0 (26 strict neighbours)
    assert np.isclose(rho.max(), 799.978, rtol=2e-6)
    assert np.isclose(rho.min(), 485.3, rtol=5e-4)
    alpha = by_uid(c, c.particle_dfsph_alphas)[mat == 1]
    # mode of alpha = interior value
    vals, cnt = np.unique(np.round(alpha, 9), return_counts=True)
    assert np.isclose(vals[np.argmax(cnt)], 1.47114e-3, rtol=2e-5)


def test_box_volumes(c1):
    c, _ = c1
    mat = by_uid(c, c.particle_materials)
    V = by_uid(c, c.particle_rest_volumes)[mat == 2] / c.V0
    assert 1.2515 < V.min() < 1.2525 and 2.0655 < V.max() < 2.0665
    vals, cnt = np.unique(np.round(V, 4), return_counts=True)
    order = np.argsort(-cnt)
    assert (vals[order[0]], cnt[order[0]]) == (1.4701, 10851)
    assert (vals[order[1]], cnt[order[1]]) == (1.7842, 5166)
    m = by_uid(c, c.particle_masses)[mat == 2]
    assert np.allclose(m, 1000.0 * V * c.V0, rtol=1e-6)


def test_neighbor_counts(c1):
    c, _ = c1
    off, idx = c.neighbor_lists()
    counts = np.diff(off)
    n = c.particle_num[None]
    mat = c.particle_materials.to_numpy(n)
    fl = counts[mat == 1]
    assert fl.min() == 7 and fl.max() == 32
    assert abs(fl.mean() - 26.4) < 0.1
    assert abs(counts[mat == 2].mean() - 15.9) < 0.1


def test_pcisph_k():
    c, s = make_sim(scene("pcisph", dt=8e-4), oracle_library())
    assert np.isclose(c.pcisph_k[None], -989.284, rtol=2e-5)


def test_scene_counts_big():
    """Particle-count formulas (base_container.py:719-751) on the 1.23M and dragon-bath blocks."""
    from sph_project_b200.containers.base_container import BaseContainer, _lattice
    count = BaseContainer.compute_cube_particle_num
    class Dummy:
        dim = 3
        particle_diameter = 0.02
    assert count(Dummy, [0.09, 0.2, 0.2], [1.7, 4.0, 1.8], space=0.02) == 1231200
    assert count(Dummy, [0.3, 0.2, 0.5], [1.2, 2.8, 1.6], space=0.02) == 321750
    Dummy._box_shell = BaseContainer._box_shell
    box = BaseContainer.compute_box_particle_num
    assert box(Dummy(), [0.04] * 3, [1 - 0.08] * 3, 0.03, space=0.02) == 17829


def test_wcsph_freefall_and_conservation():
    """Physics sanity: under-dense lattice => zero pressure => free fall for the first steps."""
    c, s = make_sim(scene("wcsph"), oracle_library())
    y0 = by_uid(c, c.particle_positions)[:, 1].copy()
    mat = by_uid(c, c.particle_materials)
    s.step(5)
    assert c.fluid_particle_num[None] == 8000 and c.particle_num[None] == 25829
    p = by_uid(c, c.particle_pressures)
    assert np.all(p[mat == 1] == 0.0)
    x = by_uid(c, c.particle_positions)
    assert np.all(x[mat == 2, 1] == y0[mat == 2])         # boundary does not move
    assert np.all(x[mat == 1, 1] < y0[mat == 1])          # fluid falls
    pad = c.padding
    assert np.all(x[mat == 1] >= pad - 1e-6) and np.all(x[mat == 1] <= 1 - pad + 1e-6)
