"""The C++ oracle (grid + sort + 27-cell window) against the independent O(N^2) numpy restatement."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, by_uid, make_sim, oracle_library, scene

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bruteforce as bf  # noqa: E402


def jittered(method, seed=0, **kw):
    """Small dam-break with jittered positions and random velocities (domain-edge cells included)."""
    c, s = make_sim(scene(method, domain_end=(0.36, 0.36, 0.32), block_start=(0.08, 0.08, 0.08),
                          block_end=(0.215, 0.235, 0.215), **kw), oracle_library())
    n = c.particle_num[None]
    rng = np.random.default_rng(seed)
    x = c.particle_positions.to_numpy(n)
    mat = c.particle_materials.to_numpy(n)
    x[mat == 1] += rng.uniform(-0.006, 0.006, size=(int((mat == 1).sum()), 3)).astype(np.float32)
    v = rng.normal(0, 0.5, size=(n, 3)).astype(np.float32)
    v[mat != 1] = 0
    c.particle_positions.from_numpy(x)
    c.particle_velocities.from_numpy(v)
    c.prepare_neighborhood_search()
    s.compute_rigid_particle_volume()
    return c, s


def state(c):
    g = lambda f: by_uid(c, f).astype(np.float64)
    return dict(x=by_uid(c, c.particle_positions), v=g(c.particle_velocities), V=g(c.particle_rest_volumes),
                m=g(c.particle_masses), rho=g(c.particle_densities), p=g(c.particle_pressures),
                mat=by_uid(c, c.particle_materials), obj=by_uid(c, c.particle_object_ids))


def test_neighbor_sets_match_bruteforce():
    c, s = jittered("wcsph")
    st = state(c)
    pairs = bf.Pairs(st["x"], c.dh)
    off, idx = c.neighbor_lists()
    n = c.particle_num[None]
    uid = c.particle_uids.to_numpy(n)
    got = np.zeros((n, n), dtype=bool)
    rows = np.repeat(np.arange(n), np.diff(off))
    got[uid[rows], uid[idx]] = True
    assert np.array_equal(got, pairs.mask)


def test_density_volume_alpha():
    c, s = jittered("dfsph", dt=1e-3)
    st = state(c)
    fl = st["mat"] == 1
    pairs = bf.Pairs(st["x"], c.dh)
    Vr = bf.rigid_volume(pairs, st["obj"], st["mat"])
    assert np.allclose(st["V"][~fl], Vr[~fl], rtol=2e-6)
    s.compute_density()
    rho = by_uid(c, c.particle_densities).astype(np.float64)
    assert np.allclose(rho[fl], bf.density(pairs, st["V"], st["mat"], 1000.0)[fl], rtol=2e-6)
    s.compute_alpha()
    a = by_uid(c, c.particle_dfsph_alphas).astype(np.float64)
    assert np.allclose(a[fl], bf.dfsph_alpha(pairs, st["V"], st["mat"])[fl], rtol=2e-5)


def test_dfsph_sweeps():
    c, s = jittered("dfsph", dt=1e-3)
    s.compute_density()
    s.compute_alpha()
    st = state(c)
    fl = st["mat"] == 1
    pairs = bf.Pairs(st["x"], c.dh)
    dchg, nn = bf.density_change(pairs, st["V"], st["v"])
    s.compute_density_derivative()
    dd = by_uid(c, c.particle_densities_derivatives).astype(np.float64)
    ref = np.where(nn < 20, 0.0, np.maximum(dchg, 0.0))
    assert np.allclose(dd[fl], ref[fl], rtol=1e-4, atol=1e-4)
    s.compute_density_star()
    ds = by_uid(c, c.particle_densities_star).astype(np.float64)
    ref = np.maximum(st["rho"] / 1000.0 + 1e-3 * dchg, 1.0)
    assert np.allclose(ds[fl], ref[fl], rtol=1e-5)
    # one divergence-correction step
    s.compute_kappa_v()
    kv = by_uid(c, c.particle_dfsph_kappa_v).astype(np.float64)
    alpha = by_uid(c, c.particle_dfsph_alphas).astype(np.float64)
    assert np.allclose(kv[fl], (dd * alpha)[fl], rtol=1e-6)
    s.correct_divergence_step()
    v1 = by_uid(c, c.particle_velocities).astype(np.float64)
    dv = bf.dfsph_correction(pairs, st["V"], st["rho"], kv, st["mat"], 1000.0, 1e-3)
    assert np.allclose((v1 - st["v"])[fl], dv[fl], rtol=2e-4, atol=2e-5)
    # constant-density variant (in-place update upstream)
    s.compute_kappa()
    k = by_uid(c, c.particle_dfsph_kappa).astype(np.float64)
    assert np.allclose(k[fl], ((ds - 1.0) * alpha / 1e-3)[fl], rtol=1e-5, atol=1e-9)
    s.correct_density_error_step()
    v2 = by_uid(c, c.particle_velocities).astype(np.float64)
    dv = bf.dfsph_correction(pairs, st["V"], st["rho"], k, st["mat"], 1000.0, 1e-3)
    assert np.allclose((v2 - v1)[fl], dv[fl], rtol=2e-4, atol=2e-5)


def test_wcsph_forces():
    c, s = jittered("wcsph")
    s.compute_density()
    # compress so that pressures are non-zero
    n = c.particle_num[None]
    rho = c.particle_densities.to_numpy(n)
    c.particle_densities.from_numpy(rho * np.float32(1.3))
    s.compute_pressure()
    st = state(c)
    fl = st["mat"] == 1
    assert np.allclose(st["p"][fl], 50000.0 * ((np.maximum(st["rho"][fl], 1000.0) / 1000.0) ** 7 - 1), rtol=2e-5, atol=0.1)
    pairs = bf.Pairs(st["x"], c.dh)
    s.compute_pressure_acceleration()
    a = by_uid(c, c.particle_accelerations).astype(np.float64)
    ref = bf.pressure_acceleration(pairs, st["V"], st["m"], st["rho"], st["p"], st["mat"], 1000.0)
    scale = np.abs(ref[fl]).max()
    assert np.allclose(a[fl], ref[fl], rtol=1e-4, atol=1e-5 * scale)
    assert np.all(a[~fl] == 0)
    # gravity + surface tension + viscosity
    s.compute_gravity_acceleration()
    s.compute_surface_tension_acceleration()
    a1 = by_uid(c, c.particle_accelerations).astype(np.float64)
    ref = np.array([0, -9.81, 0]) + bf.surface_tension(pairs, st["m"], st["mat"], 0.01, 0.02)
    assert np.allclose(a1[fl], ref[fl], rtol=1e-5, atol=1e-5)
    s.compute_viscosity_acceleration_standard()
    a2 = by_uid(c, c.particle_accelerations).astype(np.float64)
    ref = bf.viscosity_standard(pairs, st["V"], st["m"], st["rho"], st["v"], st["mat"], 1000.0, 10.0, 5.0)
    scale = np.abs(ref[fl]).max()
    assert np.allclose((a2 - a1)[fl], ref[fl], rtol=2e-4, atol=1e-5 * scale)


def test_pcisph_k_bruteforce():
    k, samples, sum_g2 = bf.pcisph_k(0.04, 0.02, 8e-4, 6.4e-6)
    assert samples == 33
    assert np.isclose(sum_g2, 1.928009e13, rtol=1e-6)
    assert np.isclose(k, -989.284, rtol=1e-6)


def test_kernel_normalisation():
    """Integral of W over the support = 1 (quadrature), W(h) = 0, gradient antisymmetric."""
    h = 0.04
    n = 64
    ax = (np.arange(n) + 0.5) / n * 2 * h - h
    P = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    r = np.sqrt((P ** 2).sum(-1))
    assert abs((bf.kernel_W(r, h)).sum() * (2 * h / n) ** 3 - 1.0) < 1e-3
    assert bf.kernel_W(h, h) == 0.0
    g1 = bf.kernel_gradient(P[:100], r[:100], h)
    g2 = bf.kernel_gradient(-P[:100], r[:100], h)
    assert np.allclose(g1, -g2)
    assert np.all(bf.kernel_gradient(np.array([[1e-6, 0, 0]]), np.array([1e-6]), h) == 0)
