#!/usr/bin/env python
"""Headline benchmark: particle-steps/s of the DFSPH dam break (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--impl reference]

Workloads (`--config`, BASELINE.json `configs`; geometry in data/scenes/*.json):
  c2p_dfsph    (default, the headline) final_scene0 geometry without its mesh bodies: 1,231,200 fluid + 727,254
               boundary particles, grid 213x200x50, DFSPH, dt = 6e-4.  N > 1: weak scaling, the block and the domain
               grow along z by one 1,231,200-particle slab per GPU, Z-slabs over NCCL inside the library.
  c2_wcsph     the same geometry under WCSPH, dt = 4e-4
  c3_bath      DFSPH bath, 321,750 fluid + 216,279 boundary particles, dt = 2e-3 (dragon_bath without its mesh bodies)
  c4_buckling  PCISPH + implicit viscosity sheet, 106,400 fluid + 2,065,095 boundary particles, dt = 1e-3
  c5_dam10m    BASELINE's 8-GPU scene: 10,000,000 fluid + 957,279 boundary particles in a 6 x 6 x 8 m tank, DFSPH, the
               same scene cut into N Z-slabs (strong scaling; meant for --gpus 8, fits from 2 GPUs up)
The lattices start 20 % under-dense, so a run first pre-rolls `--settle` untimed steps into the pressurised regime
(BASELINE.md "W-pressurised"), then W warm-up steps, then times exactly K steps between two CUDA events on the stream
the library launches on.

One JSON line on stdout (rank 0).  `value` counts FLUID particle-steps/s, whole job.
  roofline     the dominant kernel's SURVEY 8(d) compulsory bytes / its event-timed duration (a second pass of K steps
               with per-launch events; the headline pass runs without them) against the measured HBM peak; the density
               kernel the metric names is reported next to it
  e2e          the same metric through host buffers: per step upload x, v from pinned memory, sph_step(1), download x, v
  cpu_baseline the CPU oracle (a restatement of the reference's algorithm; Taichi is not installable here) stepping the
               SAME pressurised state the GPU reached, all host cores, with the parity of the two after those steps
  slab_parity  (N > 1) a small scene stepped sharded and unsharded before the run: same particles, positions, iterations

`--impl reference` times that CPU restatement alone.  When a CUDA device is present a child process pre-rolls the GPU
library to the pressurised state and writes it to a file; the timed K steps are pure oracle on that state (same window
as the GPU arm).  Without a device it falls back to the early window from the initial lattice and says so.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fluid particle-steps/s (DFSPH dam-break)"
UNIT = "particle-steps/s"

# name -> (scene file, metric label, default settle steps, human description)
CONFIGS = {
    "c2p_dfsph": ("dam_break_1m_dfsph.json", METRIC, 1000, "DFSPH dam-break, final_scene0 geometry without mesh bodies, dt=6e-4"),
    "c2_wcsph": ("dam_break_1m_wcsph.json", "fluid particle-steps/s (WCSPH dam-break)", 300, "WCSPH dam-break, final_scene0 geometry without mesh bodies, dt=4e-4"),
    "c3_bath": ("bath_500k_dfsph.json", "fluid particle-steps/s (DFSPH bath)", 300, "DFSPH bath, dragon_bath geometry without mesh bodies, dt=2e-3"),
    "c4_buckling": ("buckling_pcisph_implicit.json", "fluid particle-steps/s (PCISPH + implicit viscosity)", 50,
                    "PCISPH + implicit-viscosity sheet, final_scene3 geometry without the mesh body, dt=1e-3"),
    "c5_dam10m": ("dam_break_10m_dfsph.json", "fluid particle-steps/s (DFSPH 10M dam-break)", 300,
                  "DFSPH dam-break, 6 x 6 x 8 m tank, 10 M fluid particles, dt=6e-4 (BASELINE config 5; Z-slabs over the GPUs of one box, strong scaling)"),
}

# compulsory bytes per particle per launch (each needed field read once, each result written once);
# rows = all particles of the rank (fluid + boundary), as the survey prescribes
ALGO_BYTES = {
    "kb_build": 24, "kb_pv_sweep": 24, "kb_dfsph_density_change": 40, "kb_dfsph_correct": 52, "kb_surface_tension": 44,
    "kb_viscosity": 64, "kb_pressure_accel": 48, "kb_pcisph_density_star": 40, "kb_cg_Ap": 64, "kb_cg_prepare1": 64,
    "k_rigid_volume": 24, "k_gather": 156, "k_cell_index": 20, "k_update_velocity": 36, "k_update_position": 36,
    "k_boundary": 32, "k_prep_aux": 24,
}
LIST_CONSUMERS = ("kb_pv_sweep", "kb_dfsph_density_change", "kb_dfsph_correct", "kb_surface_tension", "kb_viscosity",
                  "kb_pressure_accel", "kb_pcisph_density_star", "kb_cg_Ap", "kb_cg_prepare1")
LIST_ENTRY_BYTES = 2   # one 16-bit window slot per accepted pair


def algo_bytes(kernel, n_total):
    """SURVEY 8(d) compulsory bytes of one launch: the per-particle figure x all particles.  kb_build<.., true, true>
    is compute_density + compute_alpha on one staged window: both sweeps' bytes."""
    base = kernel.split("<")[0]
    if base not in ALGO_BYTES:
        return None
    b = ALGO_BYTES[base] * n_total
    args = kernel.replace(" ", "")
    if base == "kb_build" and args.endswith(",true,true>"):
        b += 24 * n_total
    return b


def list_bytes(kernel, n_total, n_pairs):
    """What the neighbour lists add on top (a design artefact, not in SURVEY's figure): 2 B per list word (one count +
    one 16-bit window slot per accepted pair) for the kernel that writes them, and for every sweep that streams them."""
    base = kernel.split("<")[0]
    words = LIST_ENTRY_BYTES * (n_pairs + n_total)
    if base == "kb_build":
        flags = kernel.replace(" ", "").split("<")[-1].rstrip(">").split(",")
        if flags[0] != "true":
            return 0
        return words * (1 + sum(f == "true" for f in flags[1:]))   # written once, read back by each fused sweep
    return words if base in LIST_CONSUMERS else 0


def workload_numbers(config, n_slabs=1):
    """(n_fluid, n_total, grid) of a config from the scene description alone, with the arithmetic of the containers
    (base_container.py:74-122, 770-781, 830-835: f64 arange per axis, f32 positions, shell mask) — so that both arms
    describe the workload identically without building it."""
    sc = scene_for(config, n_slabs=n_slabs)
    cfg = sc["Configuration"]
    r = cfg["particleRadius"]
    space, dh = 2 * r, 4.0 * r
    size = np.array(cfg["domainEnd"], dtype=np.float64) - np.array(cfg["domainStart"], dtype=np.float64)
    grid = [int(g) for g in np.ceil(size / dh).astype(int)]
    n_fluid = 0
    for b in sc.get("FluidBlocks", []):
        off = np.array(b["translation"], dtype=np.float64)
        start, end = np.array(b["start"]) + off, np.array(b["end"]) + off
        n_fluid += int(np.prod([len(np.arange(b["start"][i], b["end"][i], space)) for i in range(3)]))
        del start, end
    n_box = 0
    if cfg.get("addDomainBox"):
        lo = [cfg["domainStart"][i] + dh for i in range(3)]
        ext = [size[i] - 2 * dh for i in range(3)]
        total, inner = 1, 1
        for i in range(3):
            ax = np.arange(lo[i], lo[i] + ext[i], space).astype(np.float32)
            shell = (ax <= lo[i] + 0.03) | (ax >= lo[i] + ext[i] - 0.03)
            total *= len(ax)
            inner *= int((~shell).sum())
        n_box = total - inner
    return n_fluid, n_fluid + n_box, grid


def load_traffic():
    """DRAM bytes per launch per kernel from the committed `ncu --set full` capture of the 1-GPU workload
    (profiles/r02_ncu_traffic.json).  Not live: it belongs to the N = 1 row count."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    except (OSError, ValueError):
        return {}


def roofline_report(prof, n_rows, n_pairs, clock_info, hbm_peak, peak_src, profiled_ms_per_step, traffic_table=None, n_gpus=1):
    """The `roofline` object of the JSON line from one profiled pass.
    prof: {kernel name: (launches, total ms)} (sph_profile_read); n_rows: particles a launch iterates;
    n_pairs: accepted (fluid-row) neighbour pairs; hbm_peak in GB/s."""
    traffic_table = traffic_table or {}
    per_kernel_traffic = traffic_table.get("dram_bytes_per_launch", {})
    total_kernel_ms = sum(v[1] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])

    def entry(k, v):
        ab = algo_bytes(k, n_rows)
        t_s = v[1] / v[0] * 1e-3
        e = {"name": k, "launches": int(v[0]), "ms_per_launch": v[1] / v[0], "share": v[1] / total_kernel_ms,
             "algo_GBps": (ab / t_s / 1e9) if ab else None, "frac_of_hbm_peak": (ab / t_s / 1e9 / hbm_peak) if ab else None}
        lb = list_bytes(k, n_rows, n_pairs)
        if ab and lb:
            e["with_list_bytes_GBps"] = (ab + lb) / t_s / 1e9
        return e

    kernels = [entry(k, v) for k, v in top[:12]]
    dom_name, (dom_launches, dom_ms) = top[0]
    dom_bytes = algo_bytes(dom_name, n_rows) or 0
    t_dom = dom_ms / dom_launches * 1e-3
    achieved = dom_bytes / t_dom / 1e9
    traffic = per_kernel_traffic.get(dom_name.split("<")[0]) if n_gpus == 1 else None
    roofline = {"kernel": dom_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": (traffic_table.get("source") if traffic is not None else
                                   "null: the committed ncu capture is of the 1-GPU row count"),
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms / dom_launches,
                "share_of_kernel_time": dom_ms / total_kernel_ms, "accepted_pairs": n_pairs, "rows": n_rows,
                "with_list_bytes": {"bytes_per_launch": dom_bytes + list_bytes(dom_name, n_rows, n_pairs),
                                    "GBps": (dom_bytes + list_bytes(dom_name, n_rows, n_pairs)) / t_dom / 1e9,
                                    "what": "compulsory bytes + the 16-bit neighbour lists the kernel streams (a design artefact, not part of frac)"},
                "note": "neighbour sweeps are not HBM-bound (SURVEY 8(d)): the brick kernels are bound by shared-memory bank "
                        "conflicts of the random window reads (ncu: 8.2 wavefronts per LDS.128, ideal 4) and by instruction issue; "
                        "frac = SURVEY 8(d) compulsory bytes / time / measured HBM peak",
                "profiled_pass_ms_per_step": profiled_ms_per_step, "kernels": kernels}
    # the kernel BASELINE.json's metric names: compute_density (here fused with the list build and, in DFSPH, compute_alpha)
    dens = [(k, v) for k, v in prof.items() if k.split("<")[0] == "kb_build"]
    if dens:
        k, v = max(dens, key=lambda kv: kv[1][1])
        e = entry(k, v)
        e["traffic"] = per_kernel_traffic.get("kb_build") if n_gpus == 1 else None
        e["what"] = "compute_density fused with the neighbour-list build (and compute_alpha in the DFSPH step)"
        roofline["density_kernel"] = e
    # secondary figures SURVEY 8(d) asks for next to the HBM fraction of a neighbour sweep
    sm_mhz = (clock_info or {}).get("sm_mhz") or 1965.0
    if dom_name.split("<")[0] in LIST_CONSUMERS and n_pairs:
        flops = n_pairs * 25.0                                    # f_task of the DFSPH / pressure / viscosity tasks
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        roofline["fp32_pair_model"] = {"flops_per_launch": flops, "achieved_TFLOPs": flops / t_dom / 1e12, "peak_TFLOPs": fp32_peak,
                                       "frac": flops / t_dom / 1e12 / fp32_peak,
                                       "peak_source": "148 SMs x 128 FMA lanes x 2 x sampled SM clock (nominal, not measured)"}
        roofline["pairs_per_clk_per_sm"] = n_pairs / (t_dom * 148 * sm_mhz * 1e6)
    return roofline


def scene_for(config, n_slabs=1, scale=1.0):
    """Scene dict of a config.  The dam breaks follow final_scene0.json (reference data/scenes/final_scene0.json:5-16,
    52-63) without RigidBodies; n_slabs > 1 stretches block and domain along z by 1.6 m (80 particle layers = 1,231,200
    fluid particles) per extra slab (weak scaling); scale < 1 shrinks every length (time-boxed CPU sample)."""
    if config in ("c2p_dfsph", "c2_wcsph") and (n_slabs > 1 or scale != 1.0):
        method = "dfsph" if config == "c2p_dfsph" else "wcsph"
        z = (1.6 * n_slabs + 0.4) * scale
        cfg = {"domainStart": [0.0, 0.0, 0.0], "domainEnd": [8.5 * scale, 8.0 * scale, z], "particleRadius": 0.01,
               "addDomainBox": True, "density0": 1000, "gravitation": [0.0, -9.81, 0.0], "simulationMethod": method,
               "viscosityMethod": "standard", "timeStepSize": 6e-4 if method == "dfsph" else 4e-4, "viscosity": 10.0,
               "viscosity_b": 0.3, "exportFrame": False, "exportPly": False, "exportObj": False}
        block = {"objectId": 0, "start": [0.09, 0.2, 0.2], "end": [0.09 + 1.61 * scale, 0.2 + 3.8 * scale, z - 0.2],
                 "translation": [0.0, 0.0, 0.0], "scale": [1, 1, 1], "velocity": [0.0, -0.5, 0.0], "density": 1000.0,
                 "color": [50, 100, 200], "entryTime": -1.0}
        return {"Configuration": cfg, "FluidBlocks": [block]}
    sc = json.load(open(os.path.join(ROOT, "data", "scenes", CONFIGS[config][0])))
    sc["Configuration"].update({"exportFrame": False, "exportPly": False, "exportObj": False})
    return sc


def make_sim(scene_dict, lib=None, device=0, slab=None):
    from sph_project_b200.containers import DFSPHContainer, PCISPHContainer, WCSPHContainer
    from sph_project_b200.fluid_solvers import DFSPHSolver, PCISPHSolver, WCSPHSolver
    from sph_project_b200.utils import SimConfig
    import contextlib
    import copy
    cfg = SimConfig(config=copy.deepcopy(scene_dict), verbose=False)
    C, S = {"dfsph": (DFSPHContainer, DFSPHSolver), "wcsph": (WCSPHContainer, WCSPHSolver),
            "pcisph": (PCISPHContainer, PCISPHSolver)}[cfg.get_cfg("simulationMethod")]
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
    own = container.owned_mask()
    y = x[(mat == 1) & own, 1]
    sel = y <= np.median(y)
    return float(rho[(mat == 1) & own][sel].mean())


def static_config(name, n_gpus, settle, n_fluid, n_total, grid):
    """The `config` object: what the workload is, identical in both arms (what was measured goes elsewhere)."""
    what = CONFIGS[name][3]
    return {"workload": f"{name}: {what}; {n_fluid} fluid + {n_total - n_fluid} boundary particles"
                        + ((f" ({n_gpus} Z-slabs of ~1.23 M fluid particles each, weak scaling)" if name != "c5_dam10m" else f" ({n_gpus} Z-slabs of one scene)") if n_gpus > 1 else ""),
            "n_fluid": n_fluid, "n_total": n_total, "grid": grid,
            "window": f"W-pressurised: state after {settle} settle steps from the initial lattice, then warm-up and timed steps",
            "l2": "per-step working set ~0.5 GB > 126 MB L2: no flush needed", "parallelism": f"zslab{n_gpus}"}


def solver_iterations(stats, steps):
    return {"dfsph_density": stats.total_dfsph_iterations / steps, "dfsph_divergence": stats.total_dfsph_iterations_v / steps,
            "pcisph": stats.total_pcisph_iterations / steps, "cg": stats.total_cg_iterations / steps}


def fields_by_uid(container, with_material=False):
    """(x, v[, material]) of the rank's particles ordered by uid (insertion index)."""
    from sph_project_b200._native import F
    n = container.particle_num[None]
    uid = container.engine.get_field(F.UID, n)
    out = []
    for fid in (F.POSITION, F.VELOCITY) + ((F.MATERIAL,) if with_material else ()):
        a = container.engine.get_field(fid, n)
        b = np.empty_like(a)
        b[uid] = a
        out.append(b)
    return tuple(out)


def load_state(container, solver, xs, vs, mat=None):
    """Put a state by uid — positions, velocities and the materials (emitter particles are parked as rigid until they
    cross gravitationUpper, base_solver.py:651-677) — into a freshly prepared simulation and redo what its step tail
    leaves behind (sort, densities, DFSPH alpha: all functions of the positions)."""
    from sph_project_b200._native import F
    n = container.particle_num[None]
    uid = container.engine.get_field(F.UID, n)
    if mat is not None:
        container.engine.set_field(F.MATERIAL, mat[uid])
    container.engine.set_field(F.POSITION, xs[uid])
    container.engine.set_field(F.VELOCITY, vs[uid])
    container.prepare_neighborhood_search()
    solver.compute_density()
    if hasattr(solver, "compute_alpha"):
        solver.compute_alpha()


def time_cpu(solver, container, steps, warm):
    if warm:
        solver.step(warm)
    t0 = time.perf_counter()
    st = solver.step(steps)
    dt = time.perf_counter() - t0
    return container.fluid_particle_num[None] * steps / dt, dt, st


def make_state(args):
    """Child of the reference arm: pre-roll the GPU library to the pressurised state and write (x, v) by uid."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit(3)
    c, s = make_sim(scene_for(args.config))
    c.engine.step(args.settle)
    xs, vs, mat = fields_by_uid(c, with_material=True)
    np.savez(args.make_state, x=xs, v=vs, material=mat, settle=args.settle)


def run_reference(args, rank, world):
    """CPU restatement of the reference, all host threads, the arm's own workload and window."""
    if rank != 0:
        return
    lib = oracle_library()
    cores = int(os.environ.get("OMP_NUM_THREADS", host_cores()))   # the OpenMP team size actually used
    t0 = time.perf_counter()
    # the 10 M-particle scene is sampled by the 1.23 M dam break (same solver, spacing and tank depth per slab)
    sample_config = "c2p_dfsph" if args.config == "c5_dam10m" else args.config
    sc = scene_for(sample_config)
    c, s = make_sim(sc, lib)
    n_fluid, n_total = int(c.fluid_particle_num[None]), int(c.particle_num[None])
    sample = f"{'full 1-GPU workload' if sample_config == args.config else 'bounded sample: the c2p_dfsph dam break'} ({n_fluid} fluid + {n_total - n_fluid} boundary particles)"
    window = "pressurised"
    state_path = os.path.join(tempfile.gettempdir(), f"sph_b200_state_{args.config}_{args.settle}_{os.getpid()}.npz")
    try:   # the pre-roll is the GPU library's (untimed, in a child process); the timed steps below are pure oracle
        subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--make-state", state_path, "--config", sample_config,
                        "--settle", str(args.settle)], check=True, stdout=sys.stderr, stderr=sys.stderr, timeout=900)
        st = np.load(state_path)
        load_state(c, s, st["x"], st["v"], st["material"])
        sample += f", the state the GPU library reaches after {args.settle} settle steps"
    except (subprocess.SubprocessError, OSError, KeyError, ValueError):
        window = "early"
        sample += "; no CUDA device for the pre-roll: early window from the initial lattice (1+1 solver iterations per step)"
    finally:
        if os.path.exists(state_path):
            os.remove(state_path)
    if world > 1:
        sample += f"; bounded sample of the {world}-slab weak-scaling workload: one slab (the rate per particle does not depend on the slab count)"
    value, dt, st = time_cpu(s, c, args.steps, args.warmup)
    nf_w, nt_w, grid_w = workload_numbers(args.config, n_slabs=world)
    cfg = static_config(args.config, world, args.settle, nf_w, nt_w, grid_w)
    if window == "early":
        cfg["window"] = "W-early: from the initial (20 % under-dense) lattice; no CUDA device was available to produce the pressurised state"
    line = {
        "impl": "reference", "metric": CONFIGS[args.config][1], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "stats": {"mean_iterations": solver_iterations(st, args.steps), "window": window,
                  "density_error": st.dfsph_density_error, "divergence_error": st.dfsph_divergence_error},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu": cpu_model(),
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
        if args.config not in ("c2p_dfsph", "c2_wcsph", "c5_dam10m"):
            raise SystemExit("Z-slabs are defined for the dam-break configs (c2p_dfsph, c2_wcsph: weak scaling; c5_dam10m: one scene)")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    # ---- N > 1: the Z-slab path against the unsharded run on a small scene, before anything is timed ----
    slab_parity = None
    if world > 1:
        from sph_project_b200.slab import slab_parity_check
        slab_parity = slab_parity_check(rank, world, local_rank, method="wcsph" if args.config == "c2_wcsph" else "dfsph", steps=30)

    t_setup = time.perf_counter()
    # N > 1: weak scaling, the domain and the block grow along z by one 1.23M-particle slab per GPU
    container, solver = make_sim(scene_for(args.config, n_slabs=world), device=local_rank, slab=(rank, world) if world > 1 else None)
    eng = container.engine
    n_fluid, n_total = int(container.fluid_particle_num[None]), int(container.global_particle_num)
    assert (n_fluid, n_total) == tuple(workload_numbers(args.config, n_slabs=world)[:2]), "analytic particle counts differ from the container's"
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pre-roll into the pressurised regime, then warm-up ----
    early = None
    if args.settle > 23:
        eng.step(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.step(20)
        early = {"value": n_fluid * 20 / (time.perf_counter() - t0), "window": "steps 3-23 from the initial lattice (host clock)",
                 "mean_iterations": solver_iterations(st, 20)}
        eng.step(args.settle - 23)
    else:
        eng.step(args.settle)
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
    from sph_project_b200._native import F, S
    n_local = int(container.particle_num[None])
    mat_now = container.particle_materials.to_numpy(n_local)
    n_pairs = int(eng.get_field(F.NEIGHBOR_COUNT, n_local)[(mat_now == 1) & container.owned_mask()].sum())
    roofline = roofline_report(prof, n_local, n_pairs, clock_info, hbm_peak, peak_src, prof_ms / args.steps, load_traffic(), world)
    bricks = {"active": int(eng.get_scalar(S.ACTIVE_BRICKS)), "max_window_slots": int(eng.get_scalar(S.MAX_WINDOW_SLOTS)),
              "windows_over_budget": int(eng.get_scalar(S.WINDOW_OVERFLOWS)), "cells": "4x4x3", "threads": 512}

    # ---- e2e: host buffers in and out every step, through the C ABI ----
    xh = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True).numpy()
    vh = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True).numpy()
    eng.get_field_into(F.POSITION, xh)
    eng.get_field_into(F.VELOCITY, vh)
    e2e_steps = max(3, min(args.steps, 10))
    is_dfsph = solver.__class__.__name__ == "DFSPHSolver"
    barrier()
    t0 = time.perf_counter()
    launches_e2e = 0
    for _ in range(e2e_steps):
        eng.set_field(F.POSITION, xh)
        eng.set_field(F.VELOCITY, vh)
        if is_dfsph:   # a DFSPH step starts from the previous step's tail: grid, densities, alpha of the uploaded positions
            container.prepare_neighborhood_search()
            solver.compute_density()
            solver.compute_alpha()
            launches_e2e += 11
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
           "what": "pinned host x,v -> sph_set_field x2 (DFSPH: -> prepare_neighborhood_search, compute_density, compute_alpha: the step tail a DFSPH "
                   "step starts from) -> sph_step(1) -> sph_get_field x2 -> pinned host"}

    # ---- CPU baseline + parity at the benchmark state (rank 0, N=1): both step the SAME pressurised state ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        xs, vs, mats = fields_by_uid(container, with_material=True)
        lib = oracle_library()
        c2, s2 = make_sim(scene_for(args.config), lib)
        load_state(c2, s2, xs, vs, mats)
        load_state(container, solver, xs, vs, mats)    # the GPU restarts from the very same host copy
        t0 = time.perf_counter()
        one_it = solver_iterations(s2.step(1), 1)
        one = time.perf_counter() - t0
        k = int(min(max(round(12.0 / max(one, 1e-3)), 1), 20))      # about 12 s more of CPU work
        v, dt_cpu, st2 = time_cpu(s2, c2, k, 0)
        st_gpu = eng.step(1 + k)
        xg = fields_by_uid(container)[0]
        xo = fields_by_uid(c2)[0]
        it_cpu = {key: val * k / (k + 1) + one_it[key] / (k + 1) for key, val in solver_iterations(st2, k).items()}
        cpu = {"value": v, "unit": UNIT, "cores": int(os.environ.get("OMP_NUM_THREADS", host_cores())), "kind": "port", "cpu": cpu_model(),
               "sample": f"full workload ({n_fluid} fluid + {n_total - n_fluid} boundary particles), {k} timed steps after 1 warm-up step from the "
                         f"pressurised state the GPU reached ({dt_cpu:.1f} s of CPU work)",
               "parity_at_benchmark_state": {"steps": k + 1, "mean_iterations_cpu": it_cpu, "mean_iterations_gpu": solver_iterations(st_gpu, k + 1),
                                             "max_rel_position_error": float(np.abs(xg - xo).max() / np.abs(xo).max())}}

    if rank == 0:
        line = {
            "metric": CONFIGS[args.config][1], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.config == "c5_dam10m" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": static_config(args.config, world, args.settle, *workload_numbers(args.config, n_slabs=world)),
            "stats": {"mean_iterations": solver_iterations(stats, args.steps), "window": "pressurised",
                      "density_error": stats.dfsph_density_error, "divergence_error": stats.dfsph_divergence_error,
                      "lower_half_mean_density": rho_lower, "early_window": early,
                      "total_particle_steps_per_s": n_total * args.steps / (ms * 1e-3),
                      # work per step is data dependent (the reference's convergence test averages over fluid AND boundary
                      # particles, so a scene with relatively fewer boundary particles iterates longer): rate per solver iteration
                      "fluid_particle_solver_iterations_per_s": n_fluid * (stats.total_dfsph_iterations + stats.total_dfsph_iterations_v) / (ms * 1e-3),
                      "bricks": bricks},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(stats.kernel_launches),
            "clocks": clock_info, "setup_s": time.perf_counter() - t_setup,
        }
        if world > 1:
            info = eng.slab_info()
            line["slab_parity"] = slab_parity
            line["stats"]["slab"] = {"layers_rank0": [info.z_lo, info.z_hi], "owned_rank0": info.n_owned, "ghosts_rank0": info.n_ghost,
                                     "halo_refreshes_rank0": int(info.halo_calls), "bytes_sent_rank0": int(info.halo_bytes),
                                     "transport": ("NCCL send/recv + all-reduce, issued by the library on its stream" if os.environ.get("SPH_B200_NO_PEER") == "1" else
                                                   "per step: NCCL send/recv on the library's stream; inside the DFSPH loops: ghost payloads read from the "
                                                   "neighbours' memory (CUDA IPC over NVLink) inside the sweeps, error sums stored into every rank's control block")}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c2p_dfsph", choices=sorted(CONFIGS))
    ap.add_argument("--settle", type=int, default=None, help="untimed pre-roll steps into the pressurised regime (default per config)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--make-state", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.settle is None:
        args.settle = CONFIGS[args.config][2]
    if args.make_state:
        make_state(args)
        return
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
    if args.config == "c5_dam10m":
        # the dam breaks ALONG z into slabs that start empty: room for the flood (slab.py, SlabContext.capacity)
        os.environ.setdefault("SPH_B200_SLAB_SLACK", "4.0")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    try:
        run_gpu(args, rank, world, local_rank)
    except BaseException:
        # A rank that fails (e.g. SPH_E_CAPACITY) must die at once: the interpreter's teardown would wait for this
        # rank's stream, which is parked in a collective its peers will never enter, and torchrun would never see
        # the failure.  (Measured the hard way: profiles/r02_bench_8gpu_c5_pressurised_FAILED.txt.)
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
