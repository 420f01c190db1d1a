"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties
(the CPU oracle needs seconds per step at these sizes): particle conservation, the sort's invariants,
boundary containment, solver tolerances, and bit-equality of the neighbour-list kernels with the plain
window walk (SPH_B200_NO_LISTS=1) on the same inputs."""
import os

import numpy as np
import pytest

from helpers import by_uid, make_sim, scene
from sph_project_b200._native import F

pytestmark = pytest.mark.gpu


def dam_break_1m(method, dt):
    # final_scene0 geometry without its mesh bodies (BASELINE.md C2 / C2')
    return scene(method, domain_end=(8.5, 8.0, 2.0), block_start=(0.09, 0.2, 0.2), block_end=(1.7, 4.0, 1.8),
                 velocity=(0.0, -0.5, 0.0), dt=dt, viscosity=10.0, viscosity_b=0.3)


def sort_invariants(c):
    n = c.particle_num[None]
    gid = c.grid_ids.to_numpy(n)
    uid = c.particle_uids.to_numpy(n)
    assert np.all(np.diff(gid) >= 0)                                   # cell-sorted
    assert np.array_equal(np.sort(uid), np.arange(n))                  # a permutation: nothing lost or duplicated
    same = gid[1:] == gid[:-1]
    cell = c.engine.get_field(F.CELL, n)
    nx, ny = int(c.grid_num[0]), int(c.grid_num[1])
    assert np.array_equal(gid, (cell[:, 2] * ny + cell[:, 1]) * nx + cell[:, 0])   # grid id == flatten(cell(x))
    scan = c.grid_num_particles.to_numpy()
    assert scan[-1] == n and np.all(np.diff(scan) >= 0)                # checksum of per-cell counts
    return same


@pytest.mark.parametrize("method,dt", [("dfsph", 6e-4), ("wcsph", 4e-4)])
def test_dam_break_1m_properties(method, dt):
    c, s = make_sim(dam_break_1m(method, dt))
    assert c.fluid_particle_num[None] == 1231200 and c.particle_num[None] == 1231200 + 727254
    mat0 = by_uid(c, c.particle_materials)
    x0 = by_uid(c, c.particle_positions)
    stats = s.step(8)
    n = c.particle_num[None]
    assert n == 1231200 + 727254
    sort_invariants(c)
    x = by_uid(c, c.particle_positions)
    v = by_uid(c, c.particle_velocities)
    assert np.isfinite(x).all() and np.isfinite(v).all()
    assert np.array_equal(by_uid(c, c.particle_materials), mat0)
    assert np.array_equal(x[mat0 == 2], x0[mat0 == 2])                 # boundary particles never move
    pad = c.padding
    hi = np.array([8.5, 8.0, 2.0]) - pad
    assert (x[mat0 == 1] >= pad - 1e-6).all() and (x[mat0 == 1] <= hi + 1e-6).all()
    rho = by_uid(c, c.particle_densities)[mat0 == 1]
    assert 400.0 < rho.min() and rho.max() < 1100.0
    if method == "dfsph":
        assert stats.dfsph_density_error <= 1e-4 + 1e-9 and stats.dfsph_iterations >= 1
        assert stats.dfsph_divergence_error <= 0.001 * 1000.0 / dt
    # free fall of the still under-dense block: mean vertical velocity follows g t + v0 within 2 %
    t = 8 * dt
    assert abs(v[mat0 == 1, 1].mean() - (-0.5 - 9.81 * t)) < 0.02 * (0.5 + 9.81 * t)


def test_list_kernels_equal_window_walk_bitwise():
    """Neighbour lists + record gathers change the data path, not the arithmetic or its order."""
    sc = dam_break_1m("dfsph", 6e-4)
    out = []
    for no_lists in ("0", "1"):
        os.environ["SPH_B200_NO_LISTS"] = no_lists
        try:
            c, s = make_sim(sc)
        finally:
            os.environ.pop("SPH_B200_NO_LISTS", None)
        s.step(3)
        out.append((by_uid(c, c.particle_positions), by_uid(c, c.particle_velocities), by_uid(c, c.particle_densities)))
        del c, s
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_neighbor_symmetry_and_counts_at_scale():
    """j in N(i) <=> i in N(j), on a 0.5 M-particle bath (BASELINE.md C3 geometry, no mesh bodies)."""
    sc = scene("dfsph", domain_end=(5.0, 3.0, 2.0), block_start=(0.3, 0.2, 0.5), block_end=(1.2, 2.8, 1.6),
               translation=(0.2, 0.0, 0.2), velocity=(0.0, -1.0, 0.0), dt=2e-3)
    c, s = make_sim(sc)
    assert c.fluid_particle_num[None] == 321750 and c.particle_num[None] == 321750 + 216279
    s.step(2)
    off, idx = c.neighbor_lists()
    n = c.particle_num[None]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    fwd = np.sort(rows * n + idx)
    bwd = np.sort(idx.astype(np.int64) * n + rows)
    assert np.array_equal(fwd, bwd)
    counts = np.diff(off)
    assert np.array_equal(counts, c.engine.get_field(F.NEIGHBOR_COUNT, n))
    assert counts.max() < 96          # fits the list width; larger counts fall back to the window walk
