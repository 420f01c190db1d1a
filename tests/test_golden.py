"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py with the CPU
oracle — the reference itself cannot run here, see that script).  CPU: the oracle still reproduces
them.  GPU: the CUDA path matches them to the north_star tolerance (positions within 1e-4 relative)."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, by_uid, make_sim, oracle_library

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as G  # noqa: E402


def load(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(name):
    got, want = G.run_case(name, oracle_library()), load(name)
    assert np.array_equal(got["iterations"][:3], want["iterations"][:3])
    assert abs(int(got["iterations"][3]) - int(want["iterations"][3])) <= 2          # CG count: f64 sum order
    for k in ("positions", "velocities", "densities"):
        assert np.allclose(got[k], want[k], rtol=1e-6, atol=1e-7), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dfsph_dam_break_2k", "wcsph_dam_break_2k", "pcisph_dam_break_2k"])
def test_cuda_matches_golden(name):
    method, dt, steps, kw = G.CASES[name]
    c, s = make_sim(G.case_scene(method, dt, kw))
    st = s.step(steps)
    want = load(name)
    mat = by_uid(c, c.particle_materials)
    x = by_uid(c, c.particle_positions)[mat == 1]
    assert np.abs(x - want["positions"]).max() / np.abs(want["positions"]).max() < 1e-4
    it = np.array([st.total_dfsph_iterations, st.total_dfsph_iterations_v, st.total_pcisph_iterations])
    assert np.all(np.abs(it - want["iterations"][:3]) <= 1)
    v = by_uid(c, c.particle_velocities)[mat == 1]
    assert np.abs(v - want["velocities"]).max() <= 1e-2 * max(np.abs(want["velocities"]).max(), 1e-6)
