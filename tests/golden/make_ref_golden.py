"""Golden vectors from the REFERENCE'S OWN PYTHON SOURCES, executed in this container.

    python tests/golden/make_ref_golden.py [case ...]        (needs /root/reference; minutes per case)

Taichi, pybullet and trimesh cannot be installed offline, so the reference (/root/reference/SPH, imported
in place, never copied) runs on tests/golden/ref_shim: a small emulation of the Taichi API that executes
the `@ti.kernel` / `@ti.func` bodies as serial Python with f32 numpy scalars (see its docstring for what
that does and does not reproduce).  Most scenes are tiny (a few hundred particles) because every
floating-point operation is an interpreted numpy call; BASELINE.json's 8k-particle config C1 runs too, at
about three (WCSPH) to five (DFSPH) minutes per step.  TI_SHIM_FMA=1 switches the emulation's dot products to
fused multiply-adds (sensitivity study; write elsewhere with REF_GOLDEN_OUT=<dir>).

Each case writes tests/golden/ref_<case>.npz: the scene (JSON), the state after `prepare()` and after
every step in a canonical particle order (lexicographic in the insertion position, which both sides keep),
and the solver iteration counts parsed from the reference's own log lines.  tests/test_ref_golden.py
checks the CPU oracle and (on a GPU) the CUDA path against them.
"""
import contextlib
import io
import json
import os
import re
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


def scene(method, dt=1e-3, viscosity_method="standard", spacing=0.09, velocity=(0.0, -1.0, 0.0), g_upper=None,
          block_end=(0.8, 0.85, 0.8), viscosity=0.05, viscosity_b=0.02, add_box=True, domain_end=1.6, rigid=False, late_block=False, mesh_bodies=False, second_block_density=None):
    cfg = {"domainStart": [0.0, 0.0, 0.0], "domainEnd": [domain_end] * 3, "particleRadius": 0.05,
           "particleSpacing": spacing, "addDomainBox": add_box, "density0": 1000, "gravitation": [0.0, -9.81, 0.0],
           "simulationMethod": method, "viscosityMethod": viscosity_method, "timeStepSize": dt,
           "viscosity": viscosity, "viscosity_b": viscosity_b, "exportFrame": False, "exportPly": False, "exportObj": False}
    if g_upper is not None:
        cfg["gravitationUpper"] = g_upper
    block = {"objectId": 0, "start": [0.35, 0.3, 0.35], "end": list(block_end), "translation": [0.0, 0.0, 0.0],
             "scale": [1, 1, 1], "velocity": list(velocity), "density": 1000.0, "color": [50, 100, 200], "entryTime": -1.0}
    out = {"Configuration": cfg, "FluidBlocks": [block]}
    if late_block:     # a second block that enters while the run is under way (insert_object is called every step)
        out["FluidBlocks"].append({"objectId": 1, "start": [0.9, 0.5, 0.9], "end": [1.2, 0.8, 1.2], "translation": [0.0, 0.0, 0.0],
                                   "scale": [1, 1, 1], "velocity": [0.0, -2.0, 0.0], "density": 1000.0, "color": [200, 50, 50],
                                   "entryTime": 0.0008})
    if second_block_density is not None:
        # a lighter block next to the first from the start: particle masses (V0 x density, base_container.py:411) differ
        # between neighbours in the pressure / viscosity / surface-tension sums
        out["FluidBlocks"].append({"objectId": 1, "start": [0.82, 0.3, 0.35], "end": [1.1, 0.7, 0.8], "translation": [0.0, 0.0, 0.0],
                                   "scale": [1, 1, 1], "velocity": [-1.0, 0.0, 0.0], "density": second_block_density,
                                   "color": [200, 50, 50], "entryTime": -1.0})
    if mesh_bodies:
        # a static, rotated rigid cube next to the block and a fluid body cut from a mesh (base_container.py:611-717)
        out["RigidBodies"] = [{"objectId": 1, "geometryFile": "cube.obj", "translation": [1.05, 0.45, 0.6],
                               "rotationAxis": [0.0, 1.0, 0.0], "rotationAngle": 30.0, "scale": [0.3, 0.3, 0.3],
                               "velocity": [0.0, 0.0, 0.0], "density": 1000.0, "color": [200, 100, 50], "isDynamic": False,
                               "entryTime": -1.0}]
        out["FluidBodies"] = [{"objectId": 2, "geometryFile": "cube.obj", "translation": [0.6, 1.1, 0.6],
                               "rotationAxis": [0.0, 0.0, 1.0], "rotationAngle": 20.0, "scale": [0.3, 0.25, 0.3],
                               "velocity": [0.0, -1.5, 0.0], "density": 1000.0, "color": [50, 200, 50], "entryTime": -1.0}]
    if rigid:
        # a light dynamic cube (side 0.3) dropped onto the block; geometryFile is filled in per run (temp dir)
        out["RigidBodies"] = [{"objectId": 1, "geometryFile": "cube.obj", "translation": [0.575, 1.10, 0.575],
                               "rotationAxis": [0.0, 1.0, 0.0], "rotationAngle": 0.0, "scale": [0.3, 0.3, 0.3],
                               "velocity": [0.3, -1.0, 0.0], "density": 500.0, "color": [200, 100, 50], "isDynamic": True,
                               "entryTime": -1.0}]
    return out


CUBE_OBJ = """v -0.5 -0.5 -0.5
v 0.5 -0.5 -0.5
v 0.5 0.5 -0.5
v -0.5 0.5 -0.5
v -0.5 -0.5 0.5
v 0.5 -0.5 0.5
v 0.5 0.5 0.5
v -0.5 0.5 0.5
f 1 3 2
f 1 4 3
f 5 6 7
f 5 7 8
f 1 2 6
f 1 6 5
f 4 7 3
f 4 8 7
f 1 5 8
f 1 8 4
f 2 3 7
f 2 7 6
"""


CASES = {
    # name: (scene kwargs, steps).  spacing 0.08 packs the block 1.56x over rest density so that the pressure
    # solvers iterate; 0.09 is a gentle 1.1x.  domain_end 1.7 keeps the 0.08-spaced box out of the last cell layer,
    # whose neighbour walk reads grid_num_particles out of bounds upstream (base_container.py:553-557)
    "dfsph": (dict(method="dfsph", spacing=0.08, domain_end=1.7), 3),
    "dfsph_gentle": (dict(method="dfsph"), 3),
    "wcsph": (dict(method="wcsph", dt=5e-4), 4),
    "pcisph": (dict(method="pcisph", dt=1e-3, spacing=0.08, domain_end=1.7), 3),
    "dfsph_implicit": (dict(method="dfsph", viscosity_method="implicit", viscosity=50.0, viscosity_b=20.0), 2),
    # no domain box: its particles carry object id -1 after init_object_id (base_solver.py:680-684), which the
    # emitter branch of update_fluid_position would use as an out-of-bounds index (base_solver.py:660-663)
    "dfsph_emitter": (dict(method="dfsph", g_upper=0.7, add_box=False), 3),
    # fluid <-> dynamic rigid body coupling; the rigid dynamics are ref_shim/free_body.py on both sides, not Bullet
    "dfsph_rigid": (dict(method="dfsph", rigid=True), 3),
    "wcsph_rigid": (dict(method="wcsph", dt=5e-4, rigid=True), 3),
    "pcisph_rigid": (dict(method="pcisph", rigid=True), 2),
    "wcsph_late_block": (dict(method="wcsph", dt=5e-4, late_block=True), 4),
    # mesh bodies: trimesh is replaced by this repository's voxeliser on both sides (ref_shim/trimesh.py), which pins
    # the placement / lattice / insertion logic around it and the static-body kernels, not the voxeliser itself
    "dfsph_mesh_bodies": (dict(method="dfsph", mesh_bodies=True), 2),
    # the solver / viscosity keys are independent (base_solver.py:25,195-200): BASELINE config C4's combination
    "pcisph_implicit": (dict(method="pcisph", viscosity_method="implicit", viscosity=50.0, viscosity_b=20.0, spacing=0.085), 2),
    "wcsph_implicit": (dict(method="wcsph", dt=5e-4, viscosity_method="implicit", viscosity=20.0, viscosity_b=20.0), 2),
    "wcsph_two_densities": (dict(method="wcsph", dt=5e-4, second_block_density=600.0), 3),
    "dfsph_two_densities": (dict(method="dfsph", second_block_density=600.0), 2),
    # BASELINE.json configs[0] ("C1", the reference's own CPU-runnable case): 8,000 fluid + 17,829 box particles at the
    # shipped resolution (r = 0.01).  ~3 min per step under the emulation; only the prepared and the final state are kept.
    "c1_wcsph_8k": ("data/scenes/dam_break_8k_wcsph.json", 3, dict(keep_steps=(3,))),
    # the same geometry under the north-star method (BASELINE's DFSPH configs use this resolution)
    "c1_dfsph_8k": ("data/scenes/dam_break_8k_wcsph.json", 2, dict(keep_steps=(2,), override={"simulationMethod": "dfsph", "timeStepSize": 1e-3})),
}

STATE_FIELDS = ("particle_positions", "particle_velocities", "particle_densities", "particle_pressures",
                "particle_accelerations", "particle_rest_volumes", "particle_masses", "particle_materials",
                "particle_object_ids", "particle_is_dynamic")
EXTRA_FIELDS = {"dfsph": ("particle_dfsph_alphas", "particle_densities_star", "particle_densities_derivatives"),
                "pcisph": ("particle_densities_star",), "wcsph": ()}


def import_reference():
    sys.path[:] = [os.path.join(HERE, "ref_shim"), REFERENCE] + [p for p in sys.path
                                                                 if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
    import taichi as ti
    assert "ref_shim" in ti.__file__
    from SPH.utils import SimConfig
    from SPH.containers import DFSPHContainer, WCSPHContainer, PCISPHContainer
    from SPH.fluid_solvers import DFSPHSolver, WCSPHSolver, PCISPHSolver
    import SPH
    assert SPH.__path__[0].startswith(REFERENCE), SPH.__path__
    return SimConfig, {"dfsph": (DFSPHContainer, DFSPHSolver), "wcsph": (WCSPHContainer, WCSPHSolver),
                       "pcisph": (PCISPHContainer, PCISPHSolver)}


def snapshot(container, method, order=None):
    n = int(container.particle_num[None])
    x0 = container.rigid_particle_original_positions.to_numpy()[:n]
    perm = np.lexsort((x0[:, 2], x0[:, 1], x0[:, 0]))
    out = {"x0": x0[perm]}
    for name in STATE_FIELDS + EXTRA_FIELDS[method]:
        out[name] = getattr(container, name).to_numpy()[:n][perm]
    # the sort structures of the last neighbourhood search (base_container.py:495-547): per-particle flat cell id
    # (z-fastest flatten) and the inclusive scan of the per-cell counts
    out["grid_ids"] = container.grid_ids.to_numpy()[:n][perm]
    out["grid_num_particles"] = container.grid_num_particles.to_numpy()
    return out


ITER_RE = {
    "dfsph": re.compile(r"DFSPH - iterations: (\d+)"), "dfsph_v": re.compile(r"DFSPH - iteration V: (\d+)"),
    "pcisph": re.compile(r"PCISPH - iteration: (\d+)"), "cg": re.compile(r"CG iteration:\s+(\d+)"),
}


def run_case(name):
    kw, steps, *rest = CASES[name]
    opts = rest[0] if rest else {}
    if isinstance(kw, str):      # a scene file of this repository (same JSON schema as the reference's)
        with open(os.path.join(os.path.dirname(os.path.dirname(HERE)), kw)) as fh:
            sc = json.load(fh)
        sc["Configuration"].update({"exportPly": False, "exportFrame": False})
        sc["Configuration"].update(opts.get("override", {}))
        method = sc["Configuration"]["simulationMethod"]
    else:
        sc = scene(**kw)
        method = kw["method"]
    SimConfig, classes = import_reference()
    meshes = sc.get("RigidBodies", []) + sc.get("FluidBodies", [])
    rigid = any(b.get("isDynamic") for b in sc.get("RigidBodies", []))
    tmpdir = tempfile.mkdtemp()
    for body in meshes:
        body["geometryFile"] = os.path.join(tmpdir, "cube.obj")
        with open(body["geometryFile"], "w") as fh:
            fh.write(CUBE_OBJ)
    with open(os.path.join(tmpdir, "scene.json"), "w") as fh:
        json.dump(sc, fh)
    log = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(log):
        cfg = SimConfig(scene_file_path=fh.name)
        C, S = classes[method]
        container = C(cfg, GGUI=False)
        solver = S(container)
        solver.prepare()
    import copy
    import shutil
    shutil.rmtree(tmpdir)
    sc_out = copy.deepcopy(sc)
    for body in sc_out.get("RigidBodies", []) + sc_out.get("FluidBodies", []):      # SimConfig keeps the dict: drop what load_rigid_body attached
        for key in ("mesh", "restPosition", "restCenterOfMass", "particleNum", "voxelizedPoints"):
            body.pop(key, None)
        body["geometryFile"] = "cube.obj"
    for blk in sc_out["FluidBlocks"]:
        blk.pop("particleNum", None)
    out = {"scene": np.array(json.dumps(sc_out)), "steps": np.array(steps)}
    if meshes:
        out["cube_obj"] = np.array(CUBE_OBJ)
    if rigid:
        import pybullet
        out["rigid_mass"] = np.array(container.rigid_body_masses[1], dtype=np.float32)
    for k, v in snapshot(container, method).items():
        out["prepared_" + k] = v
    if method == "pcisph":
        out["pcisph_k"] = np.array(container.pcisph_k[None], dtype=np.float32)
    print(f"[{name}] prepared {int(container.particle_num[None])} particles "
          f"({int(container.fluid_particle_num[None])} fluid) in {time.time() - t0:.0f} s", flush=True)
    iters = {k: [] for k in ITER_RE}
    for s in range(steps):
        log.seek(0); log.truncate()
        t0 = time.time()
        with contextlib.redirect_stdout(log):
            solver.step()
        text = log.getvalue()
        for k, rx in ITER_RE.items():
            iters[k].append(sum(int(m) for m in rx.findall(text)))
        for k, v in snapshot(container, method).items():
            if k in ("particle_object_ids", "particle_is_dynamic", "particle_masses"):
                continue
            if "keep_steps" in opts and (s + 1) not in opts["keep_steps"]:
                continue
            out[f"step{s + 1}_" + k] = v
        if rigid:
            F, T = pybullet.world.log[-1][0]
            out[f"step{s + 1}_rigid_force"], out[f"step{s + 1}_rigid_torque"] = F, T
            for key in ("centers_of_mass", "rotations", "velocities", "angular_velocities"):
                out[f"step{s + 1}_rigid_body_{key}"] = getattr(container, "rigid_body_" + key).to_numpy()[1]
        print(f"[{name}] step {s + 1}: {time.time() - t0:.0f} s  " + " ".join(f"{k}={v[-1]}" for k, v in iters.items()), flush=True)
    for k, v in iters.items():
        out["iterations_" + k] = np.array(v, dtype=np.int64)
    # REF_GOLDEN_OUT redirects the output (sensitivity runs such as TI_SHIM_FMA=1 must not overwrite the fixtures)
    np.savez_compressed(os.path.join(os.environ.get("REF_GOLDEN_OUT", HERE), f"ref_{name}.npz"), **out)


if __name__ == "__main__":
    for case in (sys.argv[1:] or list(CASES)):
        run_case(case)
