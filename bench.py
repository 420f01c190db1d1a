#!/usr/bin/env python
"""Headline benchmark: particle-steps/s of the DFSPH dam break (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 ("C2'" of BASELINE.md): final_scene0 geometry without its mesh bodies —
1,231,200 fluid + 727,254 boundary particles, grid 213x200x50, DFSPH, dt = 6e-4, standard viscosity.
The lattice starts 20 % under-dense, so the run first pre-rolls `--settle` (default 1000) untimed
steps into the pressurised regime (BASELINE.md "W-pressurised"), then W warm-up steps, then times
exactly K steps between two CUDA events on the stream the library launches on.

One JSON line on stdout (rank 0).  `value` counts FLUID particle-steps/s, whole job.
  roofline     the dominant kernel's algorithmic bytes / its event-timed duration (second pass of K
               steps with per-launch events; the headline pass runs without them)
  e2e          the same metric through host buffers: per step upload x, v from pinned memory,
               sph_step(1), download x, v
  cpu_baseline the CPU oracle (a restatement of the reference's algorithm; Taichi is not
               installable here) on the same workload for about 12 s, all host cores, early window

N > 1 (torchrun, one rank per GPU): weak scaling — domain and fluid block grow along z by one
1,231,200-particle slab per GPU, Z-slabs over NCCL inside the library (sph_project_b200/csrc/sph_slab.cu).

`--impl reference` times that CPU restatement on the full workload instead (early window: it cannot
afford the 1000-step pre-roll; its early-window rate is an upper bound of its pressurised rate).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fluid particle-steps/s (DFSPH dam-break)"
UNIT = "particle-steps/s"

# compulsory bytes per particle per launch; rows = all particles (fluid + boundary)
ALGO_BYTES = {
    "k_density": 24, "k_dfsph_alpha": 24, "k_dfsph_density_change": 40, "k_dfsph_correct": 52, "k_surface_tension": 44,
    "k_viscosity": 64, "k_pressure_accel": 48, "k_rigid_volume": 24, "k_gather": 156, "k_cell_index": 20,
    "k_update_velocity": 36, "k_update_position": 36, "k_boundary": 32,
}
LIST_CONSUMERS = ("k_dfsph_alpha", "k_dfsph_density_change", "k_dfsph_correct", "k_surface_tension", "k_viscosity",
                  "k_pressure_accel")


def algo_bytes(kernel, n_total, n_pairs):
    """Algorithmic bytes of one launch: the SURVEY 8(d) per-particle figure x all particles, plus the
    neighbour list itself (4 B per accepted pair) for the kernel that writes it (k_density<.., true>)
    and for the kernels that stream it instead of re-deriving it from positions."""
    base = kernel.split("<")[0]
    if base not in ALGO_BYTES:
        return None
    b = ALGO_BYTES[base] * n_total
    args = kernel.replace(" ", "")
    if base == "k_density" and args.endswith(",true>"):
        b += 4 * n_pairs + 4 * n_total
    if base in LIST_CONSUMERS and args.endswith("true>") and not (base == "k_dfsph_density_change" and args.endswith(",false,true>")):
        b += 4 * n_pairs
    return b


def roofline_report(prof, n_rows, n_pairs, clock_info, hbm_peak, peak_src, profiled_ms_per_step):
    """The `roofline` object of the JSON line from one profiled pass.
    prof: {kernel name: (launches, total ms)} (sph_profile_read); n_rows: particles a launch iterates;
    n_pairs: accepted (fluid-row) neighbour pairs; hbm_peak in GB/s."""
    total_kernel_ms = sum(v[1] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])
    kernels = []
    for k, v in top[:10]:
        ab = algo_bytes(k, n_rows, n_pairs)
        kernels.append({"name": k, "launches": int(v[0]), "ms_per_launch": v[1] / v[0], "share": v[1] / total_kernel_ms,
                        "algo_GBps": (ab / (v[1] / v[0] * 1e-3) / 1e9) if ab else None})
    dom_name, (dom_launches, dom_ms) = top[0]
    dom_bytes = algo_bytes(dom_name, n_rows, n_pairs) or 0
    achieved = dom_bytes / (dom_ms / dom_launches * 1e-3) / 1e9
    traffic = None   # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))["dram_bytes_per_launch"].get(dom_name.split("<")[0])
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"kernel": dom_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms / dom_launches,
                "share_of_kernel_time": dom_ms / total_kernel_ms, "accepted_pairs": n_pairs,
                "note": "list-based sweeps are bound by L1 gather throughput (ncu: l1tex 80-91 % of peak, ~1 sector per pair), not by HBM; frac = algorithmic bytes / time / measured HBM peak",
                "profiled_pass_ms_per_step": profiled_ms_per_step, "kernels": kernels}
    # secondary figures SURVEY 8(d) asks for next to the HBM fraction of a neighbour sweep: the pair-model FP32
    # rate and the L1 gather rate (one 32-byte record per accepted pair; a scattered 32-lane gather is serialised
    # per distinct 128-byte line)
    sm_mhz = (clock_info or {}).get("sm_mhz") or 1965.0
    if dom_name.split("<")[0] in LIST_CONSUMERS and n_pairs:
        t_s = dom_ms / dom_launches * 1e-3
        flops = n_pairs * 25.0                                    # f_task of the DFSPH / pressure / viscosity tasks
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        roofline["fp32_pair_model"] = {"flops_per_launch": flops, "achieved_TFLOPs": flops / t_s / 1e12, "peak_TFLOPs": fp32_peak,
                                       "frac": flops / t_s / 1e12 / fp32_peak,
                                       "peak_source": "148 SMs x 128 FMA lanes x 2 x sampled SM clock (nominal, not measured)"}
        roofline["l1_gather"] = {
            "pairs_per_clk_per_sm": n_pairs / (t_s * 148 * sm_mhz * 1e6),
            "model": {"lines_per_pair": 0.66, "cycles_per_line": [1.0, 2.07], "pairs_per_clk_bounds": [0.73, 1.52],
                      "source": "profiles/r01_l1_line_model.md (host-side model of the shipped gather: distinct 128-byte lines per "
                                "warp request) x B300_MICROARCH.md L1tex rates (1.0 cycle per line across LDGs, 2.07 inside one LDG)"},
            "what": "accepted pairs per SM clock; every pair gathers one 32-byte record, the L1 serialises a warp-wide gather per "
                    "distinct 128-byte line"}
    return roofline


def dam_break_scene(method="dfsph", scale=1.0, n_slabs=1):
    """final_scene0.json geometry (reference data/scenes/final_scene0.json:5-16,52-63), no RigidBodies.
    scale < 1 shrinks every length (bounded CPU sample); n_slabs > 1 stretches the block and the
    domain along z by 1.6 m (80 particle layers = 1,231,200 fluid particles) per extra slab (weak scaling)."""
    z = (1.6 * n_slabs + 0.4) * scale
    cfg = {
        "domainStart": [0.0, 0.0, 0.0], "domainEnd": [8.5 * scale, 8.0 * scale, z], "particleRadius": 0.01,
        "addDomainBox": True, "density0": 1000, "gravitation": [0.0, -9.81, 0.0], "simulationMethod": method,
        "viscosityMethod": "standard", "timeStepSize": 6e-4 if method == "dfsph" else 4e-4, "viscosity": 10.0,
        "viscosity_b": 0.3, "exportFrame": False, "exportPly": False, "exportObj": False,
    }
    s = scale
    end = [1.7, 4.0, 1.8] if (scale == 1.0 and n_slabs == 1) else [0.09 + 1.61 * s, 0.2 + 3.8 * s, z - 0.2]
    block = {"objectId": 0, "start": [0.09, 0.2, 0.2], "end": end,
             "translation": [0.0, 0.0, 0.0], "scale": [1, 1, 1], "velocity": [0.0, -0.5, 0.0], "density": 1000.0,
             "color": [50, 100, 200], "entryTime": -1.0}
    return {"Configuration": cfg, "FluidBlocks": [block]}


def make_sim(scene_dict, lib=None, device=0, slab=None):
    from sph_project_b200.containers import DFSPHContainer, WCSPHContainer
    from sph_project_b200.fluid_solvers import DFSPHSolver, WCSPHSolver
    from sph_project_b200.utils import SimConfig
    cfg = SimConfig(config=scene_dict, verbose=False)
    C, S = (DFSPHContainer, DFSPHSolver) if cfg.get_cfg("simulationMethod") == "dfsph" else (WCSPHContainer, WCSPHSolver)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        container = C(cfg, GGUI=False, engine_library=lib, device=device, slab=slab)
        solver = S(container)
        solver.prepare()
    return container, solver


def oracle_library():
    """CPU oracle = the `reference` / cpu_baseline arm; never on the product path."""
    import ctypes
    from sph_project_b200 import _native
    path = os.path.join(ROOT, "oracle", "_build", "libsph_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=sys.stderr)
    return _native.bind(ctypes.CDLL(path))


def host_cores():
    """Usable host threads: the affinity mask, capped by a cgroup CPU quota when the container has one
    (an OpenMP team larger than the quota only oversubscribes)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]            # cgroup v2
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())           # cgroup v1
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return n


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        self.marks = []

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        self.marks.append(time.time())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        lo, hi = (self.marks + [0, 1e18])[:2] if len(self.marks) >= 2 else (0, 1e18)
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            if not (lo <= t <= hi + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def lower_half_density(container):
    n = container.particle_num[None]
    x = container.particle_positions.to_numpy(n)
    rho = container.particle_densities.to_numpy(n)
    mat = container.particle_materials.to_numpy(n)
    y = x[mat == 1, 1]
    sel = y <= np.median(y)
    return float(rho[mat == 1][sel].mean())


def time_cpu(solver, container, steps, warm):
    solver.step(warm)
    t0 = time.perf_counter()
    st = solver.step(steps)
    dt = time.perf_counter() - t0
    return container.fluid_particle_num[None] * steps / dt, dt, st


def run_reference(args, rank, world):
    """CPU restatement of the reference, all host threads, the arm's own workload."""
    if rank != 0:
        return
    lib = oracle_library()
    cores = int(os.environ.get("OMP_NUM_THREADS", host_cores()))   # the OpenMP team size actually used
    t0 = time.perf_counter()
    sc = dam_break_scene("dfsph")
    c, s = make_sim(sc, lib)
    s.step(1)
    per_step = None
    t1 = time.perf_counter()
    s.step(1)
    per_step = time.perf_counter() - t1
    sample = "full workload (1,231,200 fluid + 727,254 boundary), early window from the initial lattice"
    if per_step * (args.steps + args.warmup) > 400:   # keep the whole run within a few minutes
        del c, s
        sc = dam_break_scene("dfsph", scale=0.5)
        c, s = make_sim(sc, lib)
        sample = "half-scale geometry (155,800 fluid + 174,750 boundary), early window; the full step exceeded the time box"
    if world > 1:
        sample += f"; bounded sample of the {world}-slab weak-scaling workload: one slab (the rate per particle does not depend on the slab count)"
    value, dt, st = time_cpu(s, c, args.steps, max(args.warmup - 2, 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DFSPH dam-break 1.23M fluid particles (final_scene0 geometry, no mesh bodies), dt=6e-4",
                   "n_fluid": int(c.fluid_particle_num[None]), "n_total": int(c.particle_num[None]),
                   "mean_iterations": [st.total_dfsph_iterations / args.steps, st.total_dfsph_iterations_v / args.steps]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu": cpu_model(),
                         "note": "C++/OpenMP restatement of the reference's Taichi kernels (oracle/); Taichi itself is not installable offline"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "setup_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    # stdout carries exactly one JSON line: park fd 1 on stderr while libraries (NCCL banner) may print
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    t_setup = time.perf_counter()
    # N > 1: weak scaling, the domain and the block grow along z by one 1.23M-particle slab per GPU
    container, solver = make_sim(dam_break_scene("dfsph", n_slabs=world), device=local_rank,
                                 slab=(rank, world) if world > 1 else None)
    eng = container.engine
    n_fluid, n_total = int(container.fluid_particle_num[None]), int(container.global_particle_num)
    n_local = int(container.particle_num[None])
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pre-roll into the pressurised regime, then warm-up ----
    early = None
    if args.settle > 0:
        eng.step(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.step(20)
        early = {"value": n_fluid * 20 / (time.perf_counter() - t0), "window": "steps 3-23 from the initial lattice",
                 "mean_iterations": [st.total_dfsph_iterations / 20, st.total_dfsph_iterations_v / 20]}
        eng.step(max(args.settle - 23, 0))
    rho_lower = lower_half_density(container)
    eng.step(args.warmup)

    clocks = ClockSampler(local_rank) if rank == 0 else None   # one nvidia-smi poller, not one per rank
    time.sleep(0.3)
    # ---- timed region: exactly K steps, device-resident state ----
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    profiler_range = os.environ.get("SPH_BENCH_CUDA_PROFILER") == "1"   # ncu --profile-from-start off
    if profiler_range:
        torch.cuda.profiler.start()
    if clocks:
        clocks.mark()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        stats = eng.step(args.steps)
        ev1.record(stream)
    barrier()
    if clocks:
        clocks.mark()
    if profiler_range:
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_fluid * args.steps / (ms * 1e-3)
    clock_info = clocks.stop() if clocks else None

    # ---- roofline pass: per-launch events on the same stream ----
    eng.profile_enable(True)
    barrier()
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(stream)
    eng.step(args.steps)
    evp1.record(stream)
    barrier()
    prof = eng.profile_read()
    eng.profile_enable(False)
    prof_ms = evp0.elapsed_time(evp1)
    from sph_project_b200._native import F as _F
    n_local = int(container.particle_num[None])
    mat_now = container.particle_materials.to_numpy(n_local)
    n_pairs = int(eng.get_field(_F.NEIGHBOR_COUNT, n_local)[(mat_now == 1) & container.owned_mask()].sum())
    n_rows = n_local   # kernels of this rank stream this rank's particles
    roofline = roofline_report(prof, n_rows, n_pairs, clock_info, hbm_peak, peak_src, prof_ms / args.steps)

    # ---- e2e: host buffers in and out every step, through the C ABI ----
    from sph_project_b200._native import F
    n_local = int(container.particle_num[None])
    xh = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True).numpy()
    vh = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True).numpy()
    eng.get_field_into(F.POSITION, xh)
    eng.get_field_into(F.VELOCITY, vh)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    launches_e2e = 0
    for _ in range(e2e_steps):
        eng.set_field(F.POSITION, xh)
        eng.set_field(F.VELOCITY, vh)
        launches_e2e += eng.step(1).kernel_launches
        eng.get_field_into(F.POSITION, xh)
        eng.get_field_into(F.VELOCITY, vh)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": n_fluid * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(xh.nbytes + vh.nbytes) * world,
           "d2h_bytes_per_step": int(xh.nbytes + vh.nbytes) * world, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "what": "pinned host x,v -> sph_set_field x2 -> sph_step(1) -> sph_get_field x2 -> pinned host"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lib = oracle_library()
        c2, s2 = make_sim(dam_break_scene("dfsph"), lib)
        s2.step(1)
        t0 = time.perf_counter()
        s2.step(1)
        one = time.perf_counter() - t0
        k = int(min(max(12.0 / max(one, 1e-3), 2), 40))      # about 12 s of CPU work
        v, dt, st2 = time_cpu(s2, c2, k, 0)
        cpu = {"value": v, "unit": UNIT, "cores": int(os.environ.get("OMP_NUM_THREADS", host_cores())), "kind": "port", "cpu": cpu_model(),
               "sample": f"full workload ({c2.fluid_particle_num[None]} fluid + {c2.particle_num[None] - c2.fluid_particle_num[None]} boundary), "
                         f"{k} steps after 2 warm-up steps from the initial lattice, early window ({dt:.1f} s of CPU work, "
                         f"{st2.total_dfsph_iterations / k:.1f}+{st2.total_dfsph_iterations_v / k:.1f} solver iterations/step)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "DFSPH dam-break 1.23M fluid particles (final_scene0 geometry, no mesh bodies), dt=6e-4",
                       "n_fluid": n_fluid, "n_total": n_total, "grid": [int(g) for g in container.grid_num],
                       "window": f"W-pressurised: timed after {args.settle} settle + {args.warmup} warm-up steps",
                       "lower_half_mean_density": rho_lower,
                       "mean_iterations": [stats.total_dfsph_iterations / args.steps, stats.total_dfsph_iterations_v / args.steps],
                       "density_error": stats.dfsph_density_error, "divergence_error": stats.dfsph_divergence_error,
                       "early_window": early, "total_particle_steps_per_s": n_total * args.steps / (ms * 1e-3),
                       # work per step is data dependent (the reference's convergence test averages over fluid AND boundary
                       # particles, so a scene with relatively fewer boundary particles iterates longer): rate per solver iteration
                       "fluid_particle_solver_iterations_per_s": n_fluid * (stats.total_dfsph_iterations + stats.total_dfsph_iterations_v) / (ms * 1e-3),
                       "l2": "per-step working set ~0.5 GB > 126 MB L2: no flush needed", "parallelism": f"zslab{world}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(stats.kernel_launches),
            "clocks": clock_info, "setup_s": time.perf_counter() - t_setup,
        }
        if world > 1:
            info = eng.slab_info()
            line["config"]["slab"] = {"layers_rank0": [info.z_lo, info.z_hi], "owned_rank0": info.n_owned, "ghosts_rank0": info.n_ghost,
                                      "halo_refreshes_rank0": int(info.halo_calls), "bytes_sent_rank0": int(info.halo_bytes),
                                      "transport": "NCCL send/recv + allreduce over NVLink, issued by the library on its stream"}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--settle", type=int, default=1000, help="untimed pre-roll steps into the pressurised regime")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # OpenMP placement for the reference (CPU) arm only, fixed before the OpenMP runtime starts: the
    # unpinned 128-thread team of that arm ran 3x slower than the same sweeps inside the GPU arm's
    # cpu_baseline leg.  Never for the GPU arm: exported to all ranks of an 8-rank run it bound every
    # rank's host threads to the same cores (~70x slower, profiles/r01_bench_8gpu_INVALID_*).
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) == 0:
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        os.environ.setdefault("OMP_PLACES", "cores")
        os.environ.setdefault("OMP_DYNAMIC", "false")
        # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm is one process that should use the host
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 or "OMP_NUM_THREADS" not in os.environ:
            os.environ["OMP_NUM_THREADS"] = str(host_cores())
    elif args.impl == "b200" and int(os.environ.get("WORLD_SIZE", "1")) == 1 and "OMP_NUM_THREADS" not in os.environ:
        # single-process GPU arm: its cpu_baseline leg should not oversubscribe a cgroup CPU quota
        try:
            if host_cores() < len(os.sched_getaffinity(0)):
                os.environ["OMP_NUM_THREADS"] = str(host_cores())
        except AttributeError:
            pass
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
