"""Z-slab (multi-GPU) parity: needs >= 2 GPUs on the box; spawns tests/slab_check.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("method,late", [("dfsph", False), ("wcsph", False), ("dfsph", True), ("wcsph", True)])
def test_two_slabs_match_single_gpu(method, late):
    """late: a block enters after 5 steps inside rank 1's slab only — one rank alone adds particles, the other must
    still issue the same collectives (rank-uniform flags), and the steps run task by task through the Python solver."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SLAB_CHECK_METHOD=method, SLAB_CHECK_STEPS="30", SLAB_CHECK_LATE_BLOCK="1" if late else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "slab_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["conserved"] and res["max_rel_position_error"] < 1e-4, res


def test_boundaries_follow_the_work():
    """A block that enters inside rank 1's slab makes that rank the busier one: with a re-balancing round every 3 sorts
    the boundary moves towards it, layer by layer, and the run still equals the unsharded one."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SLAB_CHECK_METHOD="dfsph", SLAB_CHECK_STEPS="30", SLAB_CHECK_LATE_BLOCK="1", SPH_B200_REBALANCE="3")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29532", os.path.join(ROOT, "tests", "slab_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["conserved"] and res["max_rel_position_error"] < 1e-4, res
    assert res["layers_per_rank"] != res["initial_layers_per_rank"], res          # the cut moved ...
    assert res["layers_per_rank"][0][1] == res["layers_per_rank"][1][0], res      # ... and both ranks agree on it
    assert res["layers_per_rank"][0][1] > res["initial_layers_per_rank"][0][1], res   # rank 0 took layers from the busier rank 1
