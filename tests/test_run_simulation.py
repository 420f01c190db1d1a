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
    run_driver(tmp_path, monkeypatch, "dam_break_8k_wcsph.json", rounds=45)
    out = tmp_path / "dam_break_8k_wcsph_output"
    check_ply(out / "000041" / "particle_object_0.ply", 8000)
