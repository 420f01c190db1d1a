"""Parity against the REFERENCE'S OWN PYTHON SOURCES.

tests/golden/ref_*.npz were produced by tests/golden/make_ref_golden.py: /root/reference/SPH imported in
place and stepped on a small Taichi emulation (tests/golden/ref_shim; f32 numpy scalars, serial loops).
They hold, for tiny dam-break scenes, the state after prepare() and after every step plus the solver
iteration counts from the reference's log lines.

The *_rigid cases couple the fluid to a dynamic cube.  PyBullet is replaced on both sides by the same
prescribed free-body rule (tests/golden/ref_shim/free_body.py), so they pin what the hot path owns: rigid
particle insertion and mass, the force / torque the fluid kernels accumulate per object,
renew_rigid_particle_state and the Akinci volumes of a moving body.

CPU: the oracle (the checker of every GPU parity test) reproduces them -- insertion positions and all integer
fields bit for bit, floats to a few f32 ulps of the field's scale, iteration counts exactly.
GPU: the CUDA path reproduces them to the north_star tolerance (positions within 1e-4 relative).
"""
import glob
import importlib.util
import json
import os

import numpy as np
import pytest

from helpers import ROOT, make_sim, oracle_library

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz")))
INT_FIELDS = ("particle_materials", "particle_object_ids", "particle_is_dynamic")


def load(name, tmp_path=None):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    sc = json.loads(str(g["scene"]))
    if "cube_obj" in g.files:       # the mesh travels inside the fixture
        path = os.path.join(str(tmp_path), "cube.obj")
        with open(path, "w") as fh:
            fh.write(str(g["cube_obj"]))
        for body in sc.get("RigidBodies", []) + sc.get("FluidBodies", []):
            body["geometryFile"] = path
    return g, sc


def _free_body_world():
    spec = importlib.util.spec_from_file_location("_free_body", os.path.join(ROOT, "tests", "golden", "ref_shim", "free_body.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.FreeBodyWorld


class FreeBodyAdapter:
    """The reference's PyBulletSolver call sequence (bullet_solver.py:77-122,144-167) over the fixtures'
    prescribed dynamics instead of Bullet -- the same rule the reference side was stepped with."""

    def __init__(self, container, gravity, dt):
        self.container, self.total_time, self.present_rigid_object = container, 0.0, []
        self.world = _free_body_world()(dt, gravity)
        self.ids = {}
        self.is_noop = False

    def insert_rigid_object(self):
        c = self.container
        for body in c.cfg.get_rigid_bodies():
            obj = body["objectId"]
            if obj in self.present_rigid_object or body["entryTime"] > self.total_time:
                continue
            assert body["isDynamic"] and body["rotationAngle"] == 0
            mass = float(str(np.float32(c.rigid_body_masses[obj])))      # the mass travels through the URDF text upstream
            self.ids[obj] = self.world.add_body(mass, body["translation"], None, body["velocity"])
            c.rigid_body_original_centers_of_mass[obj] = np.zeros(3, np.float32)
            c.rigid_body_centers_of_mass[obj] = body["translation"]
            c.rigid_body_rotations[obj] = np.eye(3)
            c.rigid_body_velocities[obj] = body["velocity"]
            c.rigid_body_angular_velocities[obj] = np.zeros(3, np.float32)
            self.present_rigid_object.append(obj)

    def step(self):
        c = self.container
        forces, torques = c.rigid_body_forces.to_numpy(), c.rigid_body_torques.to_numpy()
        self.last_wrench = {}
        for obj, b in self.ids.items():
            self.world.apply_force(b, forces[obj])
            self.world.apply_torque(b, torques[obj])
            self.last_wrench[obj] = (forces[obj].copy(), torques[obj].copy())
        c.rigid_body_forces.fill(0.0)
        c.rigid_body_torques.fill(0.0)
        self.world.step()
        for obj, b in self.ids.items():
            st = self.world.bodies[b]
            c.rigid_body_centers_of_mass[obj] = st["x"]
            c.rigid_body_rotations[obj] = st["R"]
            c.rigid_body_velocities[obj] = st["v"]
            c.rigid_body_angular_velocities[obj] = st["w"]


def build(sc, lib, g):
    """(container, solver) prepared like the fixture's reference run."""
    if "rigid_mass" not in g.files:          # no dynamic body: the stock rigid solver has nothing to do
        return make_sim(sc, lib)
    c, s = make_sim(sc, lib, prepare=False)
    s.rigid_solver = FreeBodyAdapter(c, s.g, s.dt[None])
    s.prepare()
    return c, s


def step_counts(s):
    """One step; (dfsph, dfsph_v, pcisph, cg) iteration counts from either step path."""
    st = s.step(1)
    if st is not None:
        return ours_counts(st)
    n = getattr(s, "last_iterations", None)
    nv = getattr(s, "last_iterations_v", None)
    is_pcisph = type(s).__name__.startswith("PCISPH")
    return (0 if is_pcisph or n is None else n[0], 0 if nv is None else nv[0], n[0] if is_pcisph and n else 0, 0)


def compare_rigid(c, s, g, k, rtol):
    F, T = s.rigid_solver.last_wrench[1]
    # the per-object wrench is a sum of large cancelling per-pair terms: its error scale is the largest force of the
    # run, not the residual of one step (torques: relative to |F| x 1 m)
    f_all = max(float(np.abs(g[f"step{j}_rigid_force"]).max()) for j in range(1, int(g["steps"]) + 1))
    scale = max(float(np.abs(g[f"step{k}_rigid_force"]).max()), 1e-3 * f_all, 1e-6)
    for ours, key in ((F, "rigid_force"), (T, "rigid_torque")):
        ref = g[f"step{k}_{key}"]
        assert float(np.abs(ours - ref).max()) / scale <= rtol, (k, key, ours, ref)
    for key in ("centers_of_mass", "rotations", "velocities", "angular_velocities"):
        ours, ref = np.asarray(getattr(c, "rigid_body_" + key)[1], dtype=np.float64), g[f"step{k}_rigid_body_{key}"]
        assert np.allclose(ours, ref, rtol=rtol, atol=rtol), (k, key, ours, ref)


def canonical(container):
    n = container.particle_num[None]
    x0 = container.rigid_particle_original_positions.to_numpy(n)
    return np.lexsort((x0[:, 2], x0[:, 1], x0[:, 0])), x0


def compare(container, g, prefix, rtol, only=None):
    if prefix + "x0" not in g.files:          # large cases keep only some steps
        return {}
    perm, x0 = canonical(container)
    n = container.particle_num[None]
    ref_x0 = g[prefix + "x0"]
    assert n == ref_x0.shape[0]
    assert np.array_equal(x0[perm], ref_x0)                      # same particles, same insertion lattice
    worst = {}
    for key in g.files:
        if not key.startswith(prefix) or key.endswith("_x0"):
            continue
        name = key[len(prefix):]
        if not name.startswith("particle_") or (only is not None and name not in only):
            continue
        ours, ref = getattr(container, name).to_numpy(n)[perm], g[key]
        if name in INT_FIELDS:
            assert np.array_equal(ours, ref), (prefix, name)
            continue
        # error scale: the field's magnitude over the whole run (PCISPH pressures are k x density residuals with a
        # huge k: their f32 rounding floor is set by the largest pressure of the run, not by a quiet step); an
        # absolute floor covers all-zero fields
        scale = max([float(np.abs(g[k2]).max()) for k2 in g.files if k2.endswith("_" + name) and k2[0] in "ps"] + [1e-6])
        err = float(np.abs(ours.astype(np.float64) - ref).max()) / scale
        worst[name] = err
        assert err <= rtol, f"{prefix}{name}: {err:.3e} of the field's scale (limit {rtol:.1e})"
    return worst


def compare_sort_structures(container, g, prefix):
    """Flat cell id per particle and the inclusive scan of the per-cell counts, bit for bit (oracle only: the CUDA
    path flattens x-fastest, which is not part of the reference's API)."""
    if prefix + "grid_ids" not in g.files:
        return
    perm, _ = canonical(container)
    n = container.particle_num[None]
    assert np.array_equal(container.grid_ids.to_numpy(n)[perm], g[prefix + "grid_ids"]), prefix
    assert np.array_equal(container.grid_num_particles.to_numpy(), g[prefix + "grid_num_particles"]), prefix


def iteration_counts(g, k):
    return tuple(int(g[f"iterations_{key}"][k]) for key in ("dfsph", "dfsph_v", "pcisph", "cg"))


def ours_counts(st):
    return (st.total_dfsph_iterations, st.total_dfsph_iterations_v, st.total_pcisph_iterations, st.total_cg_iterations)


def test_fixtures_present():
    assert {"dfsph", "wcsph", "pcisph"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_sources(name, tmp_path):
    g, sc = load(name, tmp_path)
    c, s = build(sc, oracle_library(), g)
    compare(c, g, "prepared_", rtol=2e-6)
    compare_sort_structures(c, g, "prepared_")
    if "pcisph_k" in g.files:
        assert np.isclose(c.pcisph_k[None], float(g["pcisph_k"]), rtol=2e-6)
    if "rigid_mass" in g.files:
        assert np.isclose(c.rigid_body_masses[1], float(g["rigid_mass"]), rtol=1e-6)
    for k in range(int(g["steps"])):
        counts = step_counts(s)
        assert counts[:3] == iteration_counts(g, k)[:3], f"step {k + 1}"
        assert abs(counts[3] - iteration_counts(g, k)[3]) <= 1, f"step {k + 1}: CG iterations"
        compare(c, g, f"step{k + 1}_", rtol=2e-5)
        compare_sort_structures(c, g, f"step{k + 1}_")
        if "rigid_mass" in g.files:
            compare_rigid(c, s, g, k + 1, rtol=2e-5)


NEXT_ROWS = [n for n in CASES if n.endswith("_rigid") or n.endswith("_mesh_bodies")]   # SURVEY 8(f2, f3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in CASES if n not in NEXT_ROWS])
def test_cuda_matches_reference_sources(name, tmp_path):
    g, sc = load(name, tmp_path)
    c, s = make_sim(sc)
    compare(c, g, "prepared_", rtol=1e-5, only=INT_FIELDS + ("particle_positions", "particle_velocities", "particle_densities",
                                                          "particle_rest_volumes", "particle_masses", "particle_dfsph_alphas"))
    for k in range(int(g["steps"])):
        ours, ref = step_counts(s), iteration_counts(g, k)
        assert all(abs(a - b) <= 1 for a, b in zip(ours[:3], ref[:3])), (k, ours, ref)
        compare(c, g, f"step{k + 1}_", rtol=1e-4, only=("particle_positions", "particle_materials"))
        compare(c, g, f"step{k + 1}_", rtol=1e-3, only=("particle_velocities", "particle_densities"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NEXT_ROWS)   # mesh bodies / dynamic rigid bodies: SURVEY 8(f2, f3), a17
def test_cuda_rigid_coupling_matches_reference_sources(name, tmp_path):
    g, sc = load(name, tmp_path)
    c, s = build(sc, None, g)
    compare(c, g, "prepared_", rtol=1e-5, only=INT_FIELDS + ("particle_positions", "particle_velocities", "particle_densities",
                                                          "particle_rest_volumes", "particle_masses"))
    for k in range(int(g["steps"])):
        counts, ref = step_counts(s), iteration_counts(g, k)
        assert all(abs(a - b) <= 1 for a, b in zip(counts[:3], ref[:3])), (k, counts, ref)
        compare(c, g, f"step{k + 1}_", rtol=1e-4, only=("particle_positions", "particle_materials"))
        compare(c, g, f"step{k + 1}_", rtol=1e-3, only=("particle_velocities", "particle_densities"))
        if "rigid_mass" in g.files:
            compare_rigid(c, s, g, k + 1, rtol=1e-3)
