"""The drop-in driver (reference run_simulation.py): CLI, loop and on-disk PLY layout.  On CPU the
library handle is swapped for the oracle (test-only monkeypatch) so the whole script runs here."""
import os
import runpy
import sys

import numpy as np
import pytest

from helpers import ROOT, oracle_library


def run_driver(tmp_path, monkeypatch, scene_name, rounds, lib=None):
    from sph_project_b200 import _native
    if lib is not None:
        monkeypatch.setattr(_native, "_cuda_lib", lib)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["run_simulation.py", "--scene_file", os.path.join(ROOT, "data", "scenes", scene_name),
                                      "--max_rounds", str(rounds)])
    runpy.run_path(os.path.join(ROOT, "run_simulation.py"), run_name="__main__")


def check_ply(path, n_expected):
    lines = open(path).read().splitlines()
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0"
    hdr_end = lines.index("end_header")
    assert f"element vertex {n_expected}" in lines[:hdr_end]
    assert lines[hdr_end - 3:hdr_end] == ["property float x", "property float y", "property float z"]
    body = np.array([[float(t) for t in l.split()] for l in lines[hdr_end + 1:]])
    assert body.shape == (n_expected, 3) and np.isfinite(body).all()
    return body


def test_driver_writes_reference_layout(tmp_path, monkeypatch):
    run_driver(tmp_path, monkeypatch, "dam_break_8k_wcsph.json", rounds=45, lib=oracle_library())
    out = tmp_path / "dam_break_8k_wcsph_output"
    # fps 60, dt 4e-4 -> output_interval = int((1/60)/4e-4) = 41: frames 000000 and 000041
    frames = sorted(p.name for p in out.iterdir())
    assert frames == ["000000", "000041"]
    a = check_ply(out / "000000" / "particle_object_0.ply", 8000)
    b = check_ply(out / "000041" / "particle_object_0.ply", 8000)
    assert b[:, 1].mean() < a[:, 1].mean()          # the block falls between the two frames


@pytest.mark.gpu
def test_driver_on_gpu(tmp_path, monkeypatch):
    """The driver on the CUDA library writes the positions the oracle-driven driver writes (north_star: 1e-4 relative).
    The two libraries sort particles differently, so the files are compared per coordinate as sorted samples plus the
    centre of mass."""
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    run_driver(tmp_path / "gpu", monkeypatch, "dam_break_8k_wcsph.json", rounds=45)
    run_driver(tmp_path / "cpu", monkeypatch, "dam_break_8k_wcsph.json", rounds=45, lib=oracle_library())
    for frame in ("000000", "000041"):
        a = check_ply(tmp_path / "gpu" / "dam_break_8k_wcsph_output" / frame / "particle_object_0.ply", 8000)
        b = check_ply(tmp_path / "cpu" / "dam_break_8k_wcsph_output" / frame / "particle_object_0.ply", 8000)
        scale = np.abs(b).max()
        for k in range(3):
            assert np.abs(np.sort(a[:, k]) - np.sort(b[:, k])).max() < 1e-4 * scale, (frame, k)
        assert np.abs(a.mean(0) - b.mean(0)).max() < 1e-5 * scale


CUBE_OBJ = """v -0.5 -0.5 -0.5\nv 0.5 -0.5 -0.5\nv 0.5 0.5 -0.5\nv -0.5 0.5 -0.5\nv -0.5 -0.5 0.5\nv 0.5 -0.5 0.5\nv 0.5 0.5 0.5\nv -0.5 0.5 0.5
f 1 3 2\nf 1 4 3\nf 5 6 7\nf 5 7 8\nf 1 2 6\nf 1 6 5\nf 4 7 3\nf 4 8 7\nf 1 5 8\nf 1 8 4\nf 2 3 7\nf 2 7 6\n"""


def test_driver_exports_rigid_meshes(tmp_path, monkeypatch):
    """exportObj: one mesh_object_{id}.obj per rigid body and output frame, following the body
    (run_simulation.py:146-150 upstream; base_solver.py:634-640 keeps the mesh in step)."""
    import json
    from helpers import scene
    (tmp_path / "cube.obj").write_text(CUBE_OBJ)
    sc = scene("wcsph", domain_end=(0.8, 0.8, 0.8), block_start=(0.2, 0.1, 0.2), block_end=(0.5, 0.3, 0.5), dt=4e-4)
    sc["Configuration"].update({"exportPly": True, "exportObj": True, "outputInterval": 3})
    sc["RigidBodies"] = [{"objectId": 1, "geometryFile": str(tmp_path / "cube.obj"), "translation": [0.35, 0.55, 0.35],
                          "rotationAxis": [0.0, 1.0, 0.0], "rotationAngle": 0.0, "scale": [0.12, 0.12, 0.12],
                          "velocity": [0.0, -1.0, 0.0], "density": 500.0, "color": [200, 100, 50], "isDynamic": True,
                          "entryTime": -1.0}]
    (tmp_path / "cube_drop.json").write_text(json.dumps(sc))
    from sph_project_b200 import _native
    monkeypatch.setattr(_native, "_cuda_lib", oracle_library())
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["run_simulation.py", "--scene_file", str(tmp_path / "cube_drop.json"), "--max_rounds", "7"])
    runpy.run_path(os.path.join(ROOT, "run_simulation.py"), run_name="__main__")
    out = tmp_path / "cube_drop_output"
    assert sorted(p.name for p in out.iterdir()) == ["000000", "000003", "000006"]

    def vertices(frame):
        lines = (out / frame / "mesh_object_1.obj").read_text().splitlines()
        v = np.array([[float(t) for t in l.split()[1:]] for l in lines if l.startswith("v ")])
        assert v.shape == (8, 3) and sum(l.startswith("f ") for l in lines) == 12
        return v

    a, b = vertices("000000"), vertices("000006")
    assert np.allclose(a.max(0) - a.min(0), 0.12, atol=1e-6)          # scaled, placed by the rigid solver
    assert -0.01 < (b - a)[:, 1].mean() + 6 * 4e-4 * 1.0 < 0.001      # fell ~6 steps at ~1 m/s (+ gravity)
    assert (out / "000006" / "particle_object_0.ply").exists()
