"""bench.py's report plumbing on CPU: the roofline object from a recorded kernel profile, the algorithmic-bytes
model, the cgroup-aware core count, and a static check that the GPU arm references no undefined name (the arm
itself needs a B200)."""
import ast
import json
import os
import sys

import pytest

from helpers import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def recorded():
    return json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final_1gpu.json")))


def test_roofline_report_reproduces_the_recorded_line():
    d = recorded()
    r = d["roofline"]
    prof = {k["name"]: (k["launches"], k["ms_per_launch"] * k["launches"]) for k in r["kernels"]}
    out = bench.roofline_report(prof, d["config"]["n_total"], r["accepted_pairs"], d["clocks"], r["peak"], r["peak_source"],
                                r["profiled_pass_ms_per_step"])
    assert out["kernel"] == r["kernel"] and out["bound"] == "hbm" and out["unit"] == "GB/s"
    assert out["achieved"] == pytest.approx(r["achieved"], rel=1e-9)
    assert out["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert out["traffic"] == r["traffic"]
    lo, hi = out["l1_gather"]["model"]["pairs_per_clk_bounds"]
    assert lo < out["l1_gather"]["pairs_per_clk_per_sm"] < hi          # inside the modelled L1 gather window
    assert 0.0 < out["fp32_pair_model"]["frac"] < 0.2
    json.dumps(out)


def test_algorithmic_bytes_model():
    n, pairs = 1000, 30000
    assert bench.algo_bytes("k_dfsph_correct< true>", n, pairs) == 52 * n + 4 * pairs
    assert bench.algo_bytes("k_dfsph_correct<false>", n, pairs) == 52 * n
    assert bench.algo_bytes("k_density<true, true>", n, pairs) == 24 * n + 4 * pairs + 4 * n      # writes the lists
    assert bench.algo_bytes("k_gather", n, pairs) == 156 * n
    assert bench.algo_bytes("k_unknown", n, pairs) is None


def test_host_cores_is_positive_and_bounded():
    n = bench.host_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_gpu_arm_has_no_undefined_names():
    """Every name loaded inside run_gpu / roofline_report is a local, a parameter, a module global or a builtin."""
    import builtins
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    module_names = {n.id for node in tree.body for n in ast.walk(node) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}
    module_names |= {node.name for node in tree.body if isinstance(node, (ast.FunctionDef, ast.ClassDef))}
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            module_names |= {(a.asname or a.name).split(".")[0] for a in node.names}
    for fn in (node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in ("run_gpu", "roofline_report", "run_reference", "main")):
        local = {a.arg for a in fn.args.args}
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                local.add(n.id)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                local |= {(a.asname or a.name).split(".")[0] for a in n.names}
            elif isinstance(n, (ast.FunctionDef, ast.Lambda)):
                local |= {a.arg for a in n.args.args}
                if isinstance(n, ast.FunctionDef):
                    local.add(n.name)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                local.add(n.name)
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load):
                assert n.id in local or n.id in module_names or hasattr(builtins, n.id), f"{fn.name}: undefined name {n.id!r} (line {n.lineno})"
