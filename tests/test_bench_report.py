"""bench.py's report plumbing on CPU: the roofline object from a recorded kernel profile, the algorithmic-bytes
model, the analytic particle counts both arms put into `config`, the cgroup-aware core count, and a static check that
the arms reference no undefined name (the GPU arm itself needs a B200)."""
import ast
import json
import os
import sys

import pytest

from helpers import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def recorded():
    return json.load(open(os.path.join(ROOT, "profiles", "r02_bench_1gpu.json")))


def test_roofline_report_reproduces_the_recorded_line():
    d = recorded()
    r = d["roofline"]
    prof = {k["name"]: (k["launches"], k["ms_per_launch"] * k["launches"]) for k in r["kernels"]}
    out = bench.roofline_report(prof, r["rows"], r["accepted_pairs"], d["clocks"], r["peak"], r["peak_source"],
                                r["profiled_pass_ms_per_step"], bench.load_traffic(), 1)
    assert out["kernel"] == r["kernel"] and out["bound"] == "hbm" and out["unit"] == "GB/s"
    assert out["achieved"] == pytest.approx(r["achieved"], rel=1e-9)
    assert out["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    # frac counts SURVEY 8(d)'s compulsory bytes only; the neighbour lists are reported separately
    assert out["algorithmic_bytes_per_launch"] == bench.ALGO_BYTES[out["kernel"].split("<")[0]] * r["rows"]
    assert out["with_list_bytes"]["bytes_per_launch"] > out["algorithmic_bytes_per_launch"]
    assert out["density_kernel"]["name"].startswith("kb_build")
    assert 0.0 < out["fp32_pair_model"]["frac"] < 0.3
    json.dumps(out)


def test_algorithmic_bytes_model():
    n, pairs = 1000, 30000
    assert bench.algo_bytes("kb_dfsph_correct<true, false>", n) == 52 * n
    assert bench.algo_bytes("kb_build<true, true, false>", n) == 24 * n
    assert bench.algo_bytes("kb_build<true, true, true>", n) == 48 * n            # compute_density + compute_alpha
    assert bench.algo_bytes("k_gather", n) == 156 * n
    assert bench.algo_bytes("k_unknown", n) is None
    words = 2 * (pairs + n)
    assert bench.list_bytes("kb_dfsph_correct<true, false>", n, pairs) == words
    assert bench.list_bytes("kb_build<true, true, true>", n, pairs) == 3 * words   # written, read by density, read by alpha
    assert bench.list_bytes("kb_build<false, true, false>", n, pairs) == 0
    assert bench.list_bytes("k_gather", n, pairs) == 0


def test_workload_numbers_match_the_survey():
    assert bench.workload_numbers("c2p_dfsph") == (1231200, 1231200 + 727254, [213, 200, 50])
    assert bench.workload_numbers("c2_wcsph")[:2] == (1231200, 1958454)
    assert bench.workload_numbers("c3_bath") == (321750, 321750 + 216279, [125, 75, 50])
    assert bench.workload_numbers("c4_buckling") == (106400, 106400 + 2065095, [100, 500, 200])
    nf2, nt2, grid2 = bench.workload_numbers("c2p_dfsph", n_slabs=2)
    assert nf2 == 2 * 1231200 and grid2[:2] == [213, 200]


def test_both_arms_describe_the_workload_identically():
    a = bench.static_config("c2p_dfsph", 1, 1000, *bench.workload_numbers("c2p_dfsph"))
    b = bench.static_config("c2p_dfsph", 1, 1000, *bench.workload_numbers("c2p_dfsph"))
    assert a == b and "1231200 fluid" in a["workload"] and a["window"].startswith("W-pressurised")


def test_host_cores_is_positive_and_bounded():
    n = bench.host_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_arms_have_no_undefined_names():
    """Every name loaded inside the arms is a local, a parameter, a module global or a builtin."""
    import builtins
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    module_names = {n.id for node in tree.body for n in ast.walk(node) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}
    module_names |= {node.name for node in tree.body if isinstance(node, (ast.FunctionDef, ast.ClassDef))}
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            module_names |= {(a.asname or a.name).split(".")[0] for a in node.names}
    for fn in (node for node in tree.body if isinstance(node, ast.FunctionDef)):
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


def test_state_transfer_reproduces_the_run():
    """bench.py moves a pressurised state between libraries as (x, v) by uid and redoes the step tail (sort, density,
    alpha).  On one library that must continue the run exactly."""
    import numpy as np
    from helpers import oracle_library, scene
    sc = scene("dfsph", domain_end=(0.6, 0.8, 0.6), block_start=(0.1, 0.1, 0.1), block_end=(0.3, 0.5, 0.3), dt=1e-3,
               velocity=(0.0, -1.0, 0.0))
    lib = oracle_library()
    ca, sa = bench.make_sim(sc, lib)
    sa.step(5)
    xs, vs, mats = bench.fields_by_uid(ca, with_material=True)
    cb, sb = bench.make_sim(sc, lib)
    bench.load_state(cb, sb, xs, vs, mats)
    ia, ib = sa.step(3), sb.step(3)
    xa = bench.fields_by_uid(ca)[0]
    xb = bench.fields_by_uid(cb)[0]
    assert ia.total_dfsph_iterations == ib.total_dfsph_iterations
    assert np.abs(xa - xb).max() / np.abs(xa).max() < 1e-6


def test_state_transfer_carries_parked_emitter_particles():
    """With gravitationUpper set, fluid particles above it are parked as rigid until they cross it
    (base_solver.py:651-677): the material is part of the state.  A transfer without it left those rows stuck in the
    receiving simulation (found on config C4: 2.6e-3 relative position error after 21 steps)."""
    import numpy as np
    from helpers import oracle_library, scene
    sc = scene("wcsph", domain_end=(0.6, 1.0, 0.6), block_start=(0.2, 0.3, 0.2), block_end=(0.4, 0.7, 0.4), dt=4e-4,
               velocity=(0.0, -2.0, 0.0), g_upper=0.5, add_domain_box=False)
    lib = oracle_library()
    ca, sa = bench.make_sim(sc, lib)
    sa.step(40)                                    # some parked rows have crossed the line and turned fluid by now
    xs, vs, mats = bench.fields_by_uid(ca, with_material=True)
    assert (mats == 1).any() and (mats == 2).any()
    cb, sb = bench.make_sim(sc, lib)
    assert not np.array_equal(bench.fields_by_uid(cb, with_material=True)[2], mats)   # a fresh run parks more rows
    bench.load_state(cb, sb, xs, vs, mats)
    sa.step(10), sb.step(10)
    xa, xb = bench.fields_by_uid(ca)[0], bench.fields_by_uid(cb)[0]
    assert np.abs(xa - xb).max() / np.abs(xa).max() < 1e-6


def test_slab_capacity_knob(monkeypatch):
    from sph_project_b200.slab import SlabContext
    counts = [100] * 10
    ctx = SlabContext(rank=1, world=2, dh=0.04, nz=10, ranges=[(0, 5), (5, 10)])
    monkeypatch.delenv("SPH_B200_SLAB_SLACK", raising=False)
    assert ctx.capacity(counts, extra=0) == int(600 * 1.3)            # owned layers + one ghost layer, x 1.3
    monkeypatch.setenv("SPH_B200_SLAB_SLACK", "4.0")
    assert ctx.capacity(counts, extra=0) == 1000                       # never more than the whole scene
