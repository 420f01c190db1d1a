"""Multi-GPU parity check of the Z-slab path, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/slab_check.py

Every rank steps its slab of a small dam break; rank 0 also steps the same scene unsharded on its
GPU and compares positions / velocities by uid (north_star tolerance 1e-4 relative), iteration counts
and particle conservation.  Prints one JSON line; exit code 1 on a mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from helpers import make_sim, scene
    method = os.environ.get("SLAB_CHECK_METHOD", "dfsph")
    steps = int(os.environ.get("SLAB_CHECK_STEPS", "30"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scene(method, domain_end=(0.6, 0.8, 0.4 * world + 0.4), block_start=(0.1, 0.1, 0.1),
               block_end=(0.3, 0.5, 0.4 * world + 0.3), velocity=(0.0, -1.0, 0.3), dt=1e-3 if method == "dfsph" else 4e-4)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        c, s = make_sim(sc, device=local, slab=(rank, world))
    it = [0, 0]
    for _ in range(steps):
        st = s.step()
        it[0] += st.total_dfsph_iterations
        it[1] += st.total_dfsph_iterations_v
    n = c.particle_num[None]
    own = c.owned_mask()
    payload = (c.particle_uids.to_numpy(n)[own], c.particle_positions.to_numpy(n)[own], c.particle_velocities.to_numpy(n)[own],
               it, c.engine.slab_info().halo_calls)
    gathered = [None] * world
    dist.gather_object(payload, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        with contextlib.redirect_stdout(sys.stderr):
            cr, sr = make_sim(sc, device=local)
        itr = [0, 0]
        for _ in range(steps):
            st = sr.step()
            itr[0] += st.total_dfsph_iterations
            itr[1] += st.total_dfsph_iterations_v
        nr = cr.particle_num[None]
        uid_r = cr.particle_uids.to_numpy(nr)
        xr = np.empty((nr, 3), np.float32); xr[uid_r] = cr.particle_positions.to_numpy(nr)
        vr = np.empty((nr, 3), np.float32); vr[uid_r] = cr.particle_velocities.to_numpy(nr)
        uid = np.concatenate([g[0] for g in gathered])
        x = np.concatenate([g[1] for g in gathered])
        v = np.concatenate([g[2] for g in gathered])
        conserved = uid.size == nr and np.array_equal(np.sort(uid), np.arange(nr))
        xs = np.empty_like(xr); vs = np.empty_like(vr)
        if conserved:
            xs[uid] = x; vs[uid] = v
            ex = float(np.abs(xs - xr).max() / np.abs(xr).max())
            ev = float(np.abs(vs - vr).max() / max(np.abs(vr).max(), 1e-6))
        else:
            ex = ev = float("inf")
        its = gathered[0][3]
        same_it = abs(its[0] - itr[0]) <= 1 and abs(its[1] - itr[1]) <= 1
        ok = conserved and ex < 1e-4 and ev < 1e-2 and same_it
        print(json.dumps({"slab_check": method, "world": world, "steps": steps, "particles": int(nr), "conserved": bool(conserved),
                          "max_rel_position_error": ex, "max_rel_velocity_error": ev, "iterations_slab": its,
                          "iterations_single": itr, "owned_per_rank": [int(g[0].size) for g in gathered],
                          "halo_calls_rank0": int(gathered[0][4]), "ok": bool(ok)}), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
