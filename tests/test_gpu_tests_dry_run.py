"""The `-m gpu` tests cannot run here (no GPU), but their Python logic can: with SPH_GPU_TESTS_DRY_RUN=1 the
conftest swaps the library handle for the CPU oracle, so scene construction, field access, tolerances' plumbing and
fixture handling of every GPU test are executed on each CPU run.  NOT a parity claim: the comparison is oracle vs
oracle (or oracle vs fixtures), and the one test that asserts the backend really is CUDA is expected to fail."""
import os
import subprocess
import sys

from helpers import ROOT

FILES = ["test_ref_golden.py", "test_golden.py", "test_run_simulation.py", "test_gpu_rigid.py", "test_gpu_parity.py"]


def run(extra):
    env = dict(os.environ, SPH_GPU_TESTS_DRY_RUN="1")
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + extra,
                          cwd=os.path.join(ROOT, "tests"), env=env, capture_output=True, text=True, timeout=900)


def test_gpu_test_logic_runs_on_the_oracle():
    r = run(FILES + ["--deselect", "test_gpu_parity.py::test_backend_is_cuda"])
    assert r.returncode == 0, r.stdout[-3000:]


def test_backend_check_cannot_be_fooled():
    r = run(["test_gpu_parity.py::test_backend_is_cuda"])
    assert r.returncode != 0 and "oracle-cpu" in r.stdout
