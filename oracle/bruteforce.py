"""Independent O(N^2) numpy restatement (f64) of the reference's per-particle sums.

TEST INFRASTRUCTURE ONLY — imported by tests/ to validate oracle/sph_oracle.cpp (grid, sort and
27-cell window logic included) on a few thousand particles.  It shares no code with the C++
oracle or the CUDA path: dense pair matrices instead of a grid, f64 instead of f32, formulas
written from the reference's Python (paths relative to the reference checkout).

The reference ships no golden vectors; see oracle/sph_oracle.cpp for how the oracle is pinned
(reference sources stepped on a Taichi emulation, tests/test_ref_golden.py).
"""
import numpy as np

FLUID, RIGID = 1, 2


class Pairs:
    """All ordered pairs (i, j), j != i, |x_i - x_j| < h, as dense matrices."""

    def __init__(self, x, h, dist_f32=True):
        x = np.asarray(x)
        if dist_f32:
            # same neighbour set as the f32 codes: canonical squared distance, fma chain x, y, z
            xf = x.astype(np.float32)
            d = xf[:, None, :] - xf[None, :, :]
            # emulate fmaf exactly through f64 (products of f32 are exact in f64; one rounding per fma)
            d64 = d.astype(np.float64)
            r2 = (d64[..., 0] * d64[..., 0]).astype(np.float32).astype(np.float64)
            r2 = (d64[..., 1] * d64[..., 1] + r2).astype(np.float32).astype(np.float64)
            r2 = (d64[..., 2] * d64[..., 2] + r2).astype(np.float32)
            self.mask = np.sqrt(r2) < np.float32(h)
        else:
            d = x[:, None, :] - x[None, :, :]
            self.mask = np.sqrt((d * d).sum(-1)) < h
        np.fill_diagonal(self.mask, False)
        x64 = x.astype(np.float64)
        self.R = x64[:, None, :] - x64[None, :, :]          # R[i, j] = x_i - x_j
        self.r = np.sqrt((self.R ** 2).sum(-1))
        self.h = float(h)


def kernel_W(r, h):
    """Cubic spline, base_solver.py:56-78."""
    k = 8.0 / np.pi / h ** 3
    q = r / h
    return np.where(q <= 0.5, k * (6 * q ** 3 - 6 * q ** 2 + 1), np.where(q <= 1.0, 2 * k * (1 - q) ** 3, 0.0))


def kernel_gradient(R, r, h):
    """base_solver.py:80-103; returns [..., 3]."""
    k = 6.0 * 8.0 / np.pi / h ** 3
    q = r / h
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(q <= 0.5, k * q * (3 * q - 2), -k * (1 - q) ** 2) / (r * h)
    g = np.where((r > 1e-5) & (q <= 1.0), g, 0.0)
    return R * g[..., None]


def density(p, V, mat, rho0):
    """base_solver.py:521-541 (fluid rows)."""
    W = kernel_W(p.r, p.h) * p.mask
    return rho0 * (V * kernel_W(0.0, p.h) + (W * V[None, :]).sum(1))


def rigid_volume(p, obj, mat):
    """base_solver.py:105-123 (rigid rows): same-object neighbours only."""
    same = obj[:, None] == obj[None, :]
    W = kernel_W(p.r, p.h) * p.mask * same
    return 1.0 / (kernel_W(0.0, p.h) + W.sum(1))


def dfsph_alpha(p, V, mat):
    """DFSPH.py:22-62."""
    g = -V[None, :, None] * kernel_gradient(p.R, p.r, p.h) * p.mask[..., None]
    fl = (mat == FLUID)[None, :]
    S = ((g ** 2).sum(-1) * fl).sum(1)
    G = g.sum(1)
    s = S + (G ** 2).sum(-1)
    return np.where(s > 1e-5, 1.0 / np.where(s > 1e-5, s, 1.0), 0.0)


def density_change(p, V, v):
    """sum_j V_j (v_i - v_j) . grad W_ij and the neighbour count (DFSPH.py:65-126)."""
    gw = kernel_gradient(p.R, p.r, p.h) * p.mask[..., None]
    dv = v[:, None, :] - v[None, :, :]
    return (V[None, :] * (dv * gw).sum(-1)).sum(1), p.mask.sum(1)


def pressure_acceleration(p, V, m, rho, pres, mat, rho0):
    """base_solver.py:135-187 (fluid rows, no rigid wrench)."""
    gw = kernel_gradient(p.R, p.r, p.h) * p.mask[..., None]
    pi = (pres / rho ** 2)
    fl = (mat == FLUID)[None, :]
    coef = np.where(fl, -m[None, :] * (pi[:, None] + pi[None, :]), -rho0 * V[None, :] * pi[:, None])
    return (coef[..., None] * gw).sum(1)


def surface_tension(p, m, mat, sigma, diameter):
    """base_solver.py:209-229."""
    fl = (mat == FLUID)[None, :] & p.mask
    R2 = (p.R ** 2).sum(-1)
    W = np.where(R2 > diameter ** 2, kernel_W(p.r, p.h), kernel_W(diameter, p.h))
    coef = sigma / m[:, None] * m[None, :] * W * fl
    return -(coef[..., None] * p.R).sum(1)


def viscosity_standard(p, V, m, rho, v, mat, rho0, mu, mu_b, dim=3):
    """base_solver.py:231-278 (the sum already divided by rho0)."""
    gw = kernel_gradient(p.R, p.r, p.h) * p.mask[..., None]
    vxy = ((v[:, None, :] - v[None, :, :]) * p.R).sum(-1)
    denom = p.r ** 2 + 0.01 * p.h ** 2
    fl = (mat == FLUID)[None, :]
    c_f = 2 * (dim + 2) * mu * ((m[:, None] + m[None, :]) / 2) / rho[None, :]
    c_b = 2 * (dim + 2) * mu_b * (rho0 * V[None, :]) / rho[:, None]
    coef = np.where(fl, c_f, c_b) / denom * vxy
    return (coef[..., None] * gw).sum(1) / rho0


def dfsph_correction(p, V, rho, kappa, mat, rho0, dt, eps=1e-5):
    """Velocity change of one correction step (DFSPH.py:161-202 / :245-283), no rigid wrench."""
    gw = kernel_gradient(p.R, p.r, p.h) * p.mask[..., None]
    fl = (mat == FLUID)[None, :]
    ki = kappa[:, None]
    kj = kappa[None, :]
    term_f = np.where(np.abs(ki + kj) > eps * dt, ki / rho[:, None] + kj / rho[None, :], 0.0)
    term_r = np.where(np.abs(ki) > eps * dt, ki / rho[:, None], 0.0) * np.ones_like(kj)
    coef = np.where(fl, term_f, term_r) * V[None, :] * rho0
    return -(coef[..., None] * gw).sum(1)


def pcisph_k(h, diameter, dt, V0):
    """PCISPH.py:128-151."""
    diam = diameter * 0.97
    n = int(h / diam) + 1
    ax = np.arange(-n, n + 1) * diam
    P = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    R = -P
    r = np.sqrt((R ** 2).sum(-1))
    keep = r < h
    g = kernel_gradient(R[keep], r[keep], h)
    s = g.sum(0)
    return -0.5 / (dt * V0) ** 2 / ((s ** 2).sum() + (g ** 2).sum()), int(keep.sum()), float((g ** 2).sum())
