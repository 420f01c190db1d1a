import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_sessionstart(session):
    """SPH_GPU_TESTS_DRY_RUN=1: run the `-m gpu` tests' own Python logic against the CPU oracle (library handle
    swapped here, in test infrastructure).  This proves nothing about the CUDA path -- test_backend_is_cuda fails under
    it by design -- it only keeps the GPU tests from rotting between GPU runs (tests/test_gpu_tests_dry_run.py)."""
    if os.environ.get("SPH_GPU_TESTS_DRY_RUN") == "1":
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import oracle_library
        from sph_project_b200 import _native
        _native._cuda_lib = oracle_library()
