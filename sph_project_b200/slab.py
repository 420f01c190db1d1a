"""Host-side plumbing of the Z-slab decomposition; no reference counterpart.

One process per GPU (torchrun).  Every rank builds the same scene description, keeps only the
particles whose cell layer falls into its slab, and the library does the rest over NCCL
(migration, ghost import, halo refreshes, error all-reduces: sph_project_b200/csrc/sph_slab.cu).
The host side is small on purpose:

  * `balanced_ranges`   split the grid's cell layers into `world` contiguous slabs with about the
                        same number of particles each (from a cell-layer histogram);
  * `broadcast_bytes`   ship the 128-byte NCCL unique id from rank 0 with torch.distributed
                        (works on the `nccl` and on the `gloo` backend);
  * `SlabContext`       rank / world / ranges / ownership test used by the particle container;
  * `slab_parity_check` the sharded run against the unsharded one on a small dam break (used by bench.py at N > 1 and
                        by tests/slab_check.py): same particles, same positions, same solver iterations.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


def balanced_ranges(layer_counts: Sequence[int], world: int, min_layers: int = 2) -> List[Tuple[int, int]]:
    """Contiguous [z_lo, z_hi) per rank covering all layers, particle counts as even as the layer
    granularity allows; every slab gets at least `min_layers` layers (a slab must be at least as
    thick as its two ghost imports are apart)."""
    counts = np.asarray(layer_counts, dtype=np.int64)
    nz = int(counts.size)
    if world < 1 or nz < world * min_layers:
        raise ValueError(f"cannot cut {nz} cell layers into {world} slabs of >= {min_layers} layers")
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        z = int(np.searchsorted(cum, target, side="left"))
        # closest layer boundary to the target, leaving room for the remaining slabs
        if z > 0 and abs(cum[z - 1] - target) <= abs(cum[min(z, nz)] - target):
            z -= 1
        z = max(z, cuts[-1] + min_layers)
        z = min(z, nz - (world - r) * min_layers)
        cuts.append(z)
    cuts.append(nz)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def cell_layer(z: np.ndarray, dh: float, nz: int) -> np.ndarray:
    """Cell layer of z coordinates exactly as the device computes it: trunc(f32(z) / f32(dh)), clamped."""
    cz = (np.asarray(z, dtype=np.float32) / np.float32(dh)).astype(np.int64)
    return np.clip(cz, 0, nz - 1)


def broadcast_bytes(payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    """Broadcast a byte string from `src` with torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


@dataclass
class SlabContext:
    rank: int
    world: int
    dh: float
    nz: int
    ranges: List[Tuple[int, int]] = field(default_factory=list)

    @property
    def z_lo(self) -> int:
        return self.ranges[self.rank][0]

    @property
    def z_hi(self) -> int:
        return self.ranges[self.rank][1]

    def owned_z(self, z: np.ndarray) -> np.ndarray:
        cz = cell_layer(z, self.dh, self.nz)
        return (cz >= self.z_lo) & (cz < self.z_hi)

    def owned(self, positions: np.ndarray) -> np.ndarray:
        return self.owned_z(positions[:, 2])

    def capacity(self, layer_counts: Sequence[int], slack: Optional[float] = None, extra: int = 65536) -> int:
        """Local particle capacity: owned layers + one ghost layer each side, with head-room for
        migration imbalance (SPH_B200_SLAB_SLACK, default 1.3; never more than the whole scene).  Scenes whose fluid
        floods slabs that start empty (a dam break ALONG z) need a larger factor: a rank that runs out of room fails
        with SPH_E_CAPACITY."""
        import os
        if slack is None:
            slack = float(os.environ.get("SPH_B200_SLAB_SLACK", "1.3"))
        c = np.asarray(layer_counts, dtype=np.int64)
        lo, hi = max(self.z_lo - 1, 0), min(self.z_hi + 1, self.nz)
        return min(int(c[lo:hi].sum() * slack), int(c.sum())) + extra


def slab_parity_check(rank: int, world: int, local_rank: int, method: str = "dfsph", steps: int = 30,
                      late_block: bool = False) -> Optional[dict]:
    """Every rank steps its slab of a small dam break; rank 0 also steps the same scene unsharded on its GPU and
    compares positions / velocities by uid (north_star tolerance 1e-4 relative), solver iteration counts and particle
    conservation.  Needs an initialised torch.distributed process group (one rank per GPU).  Returns the report on rank
    0 (key "ok"), None elsewhere.
    late_block: a second fluid block enters after 5 steps inside the LAST rank's slab only, so one rank alone adds
    particles mid-run and the step runs task by task through the Python solver (the hooks in the middle of `_step`)."""
    import contextlib
    import sys

    import torch.distributed as dist

    from .containers import DFSPHContainer, WCSPHContainer
    from .fluid_solvers import DFSPHSolver, WCSPHSolver
    from .utils import SimConfig

    dt = 1e-3 if method == "dfsph" else 4e-4
    scene = {
        "Configuration": {
            "domainStart": [0.0, 0.0, 0.0], "domainEnd": [0.6, 0.8, 0.4 * world + 0.4], "particleRadius": 0.01, "addDomainBox": True,
            "density0": 1000.0, "gravitation": [0.0, -9.81, 0.0], "simulationMethod": method, "viscosityMethod": "standard",
            "viscosity": 10.0, "viscosity_b": 5.0, "timeStepSize": dt, "exportFrame": False, "exportPly": False, "exportObj": False},
        "FluidBlocks": [{"objectId": 0, "start": [0.1, 0.1, 0.1], "end": [0.3, 0.5, 0.4 * world + 0.3], "translation": [0, 0, 0],
                         "scale": [1, 1, 1], "velocity": [0.0, -1.0, 0.3], "density": 1000.0, "color": [50, 100, 200], "entryTime": -1.0}]}
    if late_block:
        scene["FluidBlocks"].append({"objectId": 1, "start": [0.38, 0.1, 0.4 * world + 0.1], "end": [0.5, 0.3, 0.4 * world + 0.3],
                                     "translation": [0, 0, 0], "scale": [1, 1, 1], "velocity": [0.0, -0.5, 0.0], "density": 1000.0,
                                     "color": [200, 100, 50], "entryTime": 5.5 * dt})
    C, S = (DFSPHContainer, DFSPHSolver) if method == "dfsph" else (WCSPHContainer, WCSPHSolver)

    def build(slab):
        import copy
        with contextlib.redirect_stdout(sys.stderr):
            c = C(SimConfig(config=copy.deepcopy(scene), verbose=False), GGUI=False, device=local_rank, slab=slab)
            s = S(c)
            s.prepare()
        return c, s

    def run(s):
        it = [0, 0]
        for _ in range(steps):
            st = s.step()
            if st is not None:          # native step
                it[0] += st.total_dfsph_iterations
                it[1] += st.total_dfsph_iterations_v
            elif method == "dfsph":     # task-by-task step (objects pending): the Python loops keep the counts
                it[0] += s.last_iterations[0]
                it[1] += s.last_iterations_v[0]
        return it

    c, s = build((rank, world))
    it = run(s)
    n = c.particle_num[None]
    own = c.owned_mask()
    info = c.engine.slab_info()
    payload = (c.particle_uids.to_numpy(n)[own], c.particle_positions.to_numpy(n)[own], c.particle_velocities.to_numpy(n)[own],
               it, int(info.halo_calls), (int(info.z_lo), int(info.z_hi)), tuple(int(v) for v in c.slab.ranges[rank]))
    gathered = [None] * world
    dist.gather_object(payload, gathered if rank == 0 else None, dst=0)
    del c, s
    if rank != 0:
        return None
    cr, sr = build(None)
    itr = run(sr)
    nr = cr.particle_num[None]
    uid_r = cr.particle_uids.to_numpy(nr)
    xr = np.empty((nr, 3), np.float32)
    xr[uid_r] = cr.particle_positions.to_numpy(nr)
    vr = np.empty((nr, 3), np.float32)
    vr[uid_r] = cr.particle_velocities.to_numpy(nr)
    uid = np.concatenate([g[0] for g in gathered])
    x = np.concatenate([g[1] for g in gathered])
    v = np.concatenate([g[2] for g in gathered])
    conserved = bool(uid.size == nr and np.array_equal(np.sort(uid), np.arange(nr)))
    ex = ev = float("inf")
    if conserved:
        xs, vs = np.empty_like(xr), np.empty_like(vr)
        xs[uid] = x
        vs[uid] = v
        ex = float(np.abs(xs - xr).max() / np.abs(xr).max())
        ev = float(np.abs(vs - vr).max() / max(np.abs(vr).max(), 1e-6))
    its = gathered[0][3]
    same_it = abs(its[0] - itr[0]) <= 1 and abs(its[1] - itr[1]) <= 1
    ok = bool(conserved and ex < 1e-4 and ev < 1e-2 and same_it)
    return {"method": method, "world": world, "steps": steps, "late_block": late_block, "particles": int(nr), "conserved": conserved,
            "max_rel_position_error": ex, "max_rel_velocity_error": ev, "iterations_slab": [int(a) for a in its],
            "iterations_single": [int(a) for a in itr], "owned_per_rank": [int(g[0].size) for g in gathered],
            "halo_calls_rank0": gathered[0][4], "layers_per_rank": [list(g[5]) for g in gathered],
            "initial_layers_per_rank": [list(g[6]) for g in gathered], "ok": ok}
