"""Parity against the REFERENCE'S OWN PYTHON SOURCES.

tests/golden/ref_*.npz were produced by tests/golden/make_ref_golden.py: /root/reference/SPH imported in
place and stepped on a small Taichi emulation (tests/golden/ref_shim; f32 numpy scalars, serial loops).
They hold, for tiny dam-break scenes, the state after prepare() and after every step plus the solver
iteration counts from the reference's log lines.

CPU: the oracle (the checker of every GPU parity test) reproduces them -- insertion positions and all integer
fields bit for bit, floats to a few f32 ulps of the field's scale, iteration counts exactly.
GPU: the CUDA path reproduces them to the north_star tolerance (positions within 1e-4 relative).
"""
import glob
import json
import os

import numpy as np
import pytest

from helpers import ROOT, make_sim, oracle_library

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz")))
INT_FIELDS = ("particle_materials", "particle_object_ids", "particle_is_dynamic")


def load(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    return g, json.loads(str(g["scene"]))


def canonical(container):
    n = container.particle_num[None]
    x0 = container.rigid_particle_original_positions.to_numpy(n)
    return np.lexsort((x0[:, 2], x0[:, 1], x0[:, 0])), x0


def compare(container, g, prefix, rtol, only=None):
    perm, x0 = canonical(container)
    n = container.particle_num[None]
    assert n == g["prepared_x0"].shape[0]
    assert np.array_equal(x0[perm], g["prepared_x0"])            # same particles, same insertion lattice
    worst = {}
    for key in g.files:
        if not key.startswith(prefix) or key.endswith("_x0"):
            continue
        name = key[len(prefix):]
        if only is not None and name not in only:
            continue
        ours, ref = getattr(container, name).to_numpy(n)[perm], g[key]
        if name in INT_FIELDS:
            assert np.array_equal(ours, ref), (prefix, name)
            continue
        scale = max(float(np.abs(ref).max()), 1e-6)      # all-zero fields (e.g. pressures after prepare): absolute
        err = float(np.abs(ours.astype(np.float64) - ref).max()) / scale
        worst[name] = err
        assert err <= rtol, f"{prefix}{name}: {err:.3e} of the field's scale (limit {rtol:.1e})"
    return worst


def iteration_counts(g, k):
    return tuple(int(g[f"iterations_{key}"][k]) for key in ("dfsph", "dfsph_v", "pcisph", "cg"))


def ours_counts(st):
    return (st.total_dfsph_iterations, st.total_dfsph_iterations_v, st.total_pcisph_iterations, st.total_cg_iterations)


def test_fixtures_present():
    assert {"dfsph", "wcsph", "pcisph"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_sources(name):
    g, sc = load(name)
    c, s = make_sim(sc, oracle_library())
    compare(c, g, "prepared_", rtol=2e-6)
    if "pcisph_k" in g.files:
        assert np.isclose(c.pcisph_k[None], float(g["pcisph_k"]), rtol=2e-6)
    for k in range(int(g["steps"])):
        st = s.step(1)
        assert ours_counts(st)[:3] == iteration_counts(g, k)[:3], f"step {k + 1}"
        assert abs(ours_counts(st)[3] - iteration_counts(g, k)[3]) <= 1, f"step {k + 1}: CG iterations"
        compare(c, g, f"step{k + 1}_", rtol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_sources(name):
    g, sc = load(name)
    c, s = make_sim(sc)
    compare(c, g, "prepared_", rtol=1e-5, only=INT_FIELDS + ("particle_positions", "particle_velocities", "particle_densities",
                                                          "particle_rest_volumes", "particle_masses", "particle_dfsph_alphas"))
    for k in range(int(g["steps"])):
        st = s.step(1)
        ours, ref = ours_counts(st), iteration_counts(g, k)
        assert all(abs(a - b) <= 1 for a, b in zip(ours[:3], ref[:3])), (k, ours, ref)
        compare(c, g, f"step{k + 1}_", rtol=1e-4, only=("particle_positions", "particle_materials"))
        compare(c, g, f"step{k + 1}_", rtol=1e-3, only=("particle_velocities", "particle_densities"))
