#!/usr/bin/env python
"""Step any scene of data/scenes on cuda:0 and print one JSON line: device-timed ms/step, solver iterations and the
per-kernel CUDA-event profile of a second pass.  Also the target of the ncu captures:

    ncu --profile-from-start off --set full -k regex:'k_density|k_gather' -c 4 -o gpurun_out/x \
        python tools/scene_step.py --scene data/scenes/dam_break_1m_wcsph.json --settle 200 --steps 1 --cuda-profiler

(`--cuda-profiler` brackets the timed steps with cudaProfilerStart/Stop; a time printed under ncu is not a bench value.)
"""
import argparse
import contextlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", required=True)
    ap.add_argument("--settle", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cuda-profiler", action="store_true")
    ap.add_argument("--no-profile-pass", action="store_true")
    args = ap.parse_args()
    import torch
    from sph_project_b200.containers import DFSPHContainer, PCISPHContainer, WCSPHContainer
    from sph_project_b200.fluid_solvers import DFSPHSolver, PCISPHSolver, WCSPHSolver
    from sph_project_b200.utils import SimConfig
    cfg = SimConfig(args.scene, verbose=False)
    C, S = {"wcsph": (WCSPHContainer, WCSPHSolver), "pcisph": (PCISPHContainer, PCISPHSolver),
            "dfsph": (DFSPHContainer, DFSPHSolver)}[cfg.get_cfg("simulationMethod")]
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        c = C(cfg, GGUI=False)
        s = S(c)
        s.prepare()
    eng = c.engine
    setup_s = time.perf_counter() - t0
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    if args.settle:
        eng.step(args.settle)
    torch.cuda.synchronize()
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    st = eng.step(args.steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    out = {"scene": os.path.basename(args.scene), "method": cfg.get_cfg("simulationMethod"),
           "viscosity_method": cfg.get_cfg("viscosityMethod"), "n_fluid": int(c.fluid_particle_num[None]),
           "n_total": int(c.particle_num[None]), "grid": [int(g) for g in c.grid_num], "settle": args.settle,
           "steps": args.steps, "ms_per_step": ms, "fluid_particle_steps_per_s": c.fluid_particle_num[None] / (ms * 1e-3),
           "stats": {k: (v / args.steps if k.startswith("total_") or k == "kernel_launches" else v) for k, v in st.as_dict().items()},
           "setup_s": setup_s}
    if not args.no_profile_pass:
        eng.profile_enable(True)
        eng.step(args.steps)
        prof = eng.profile_read()
        eng.profile_enable(False)
        tot = sum(v[1] for v in prof.values())
        out["kernels"] = [{"name": k, "launches_per_step": v[0] / args.steps, "ms_per_launch": v[1] / v[0], "share": v[1] / tot}
                          for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]]
        out["kernel_ms_per_step"] = tot / args.steps
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
