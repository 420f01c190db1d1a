#!/usr/bin/env python
"""Sweep-by-sweep comparison of the brick-list path with the global window walk (SPH_B200_NO_LISTS=1) on cuda:0:
both run the same kernels on the same inputs, so every field must agree bit for bit after every sweep.
Prints, per stage, the number of particles that differ and the largest difference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import contextlib
    from helpers import make_sim, scene
    from sph_project_b200._native import F, S
    size = sys.argv[1] if len(sys.argv) > 1 else "small"
    if size == "small":
        sc = scene("dfsph", domain_end=(1.0, 1.0, 1.0), block_start=(0.1, 0.1, 0.1), block_end=(0.495, 0.495, 0.495), dt=1e-3,
                   velocity=(0.0, -1.0, 0.0))
    else:
        sc = scene("dfsph", domain_end=(8.5, 8.0, 2.0), block_start=(0.09, 0.2, 0.2), block_end=(1.7, 4.0, 1.8),
                   velocity=(0.0, -0.5, 0.0), dt=6e-4, viscosity=10.0, viscosity_b=0.3)
    sims = []
    for no_lists in ("0", "1"):
        os.environ["SPH_B200_NO_LISTS"] = no_lists
        with contextlib.redirect_stdout(sys.stderr):
            sims.append(make_sim(sc))
        os.environ.pop("SPH_B200_NO_LISTS", None)
    (ca, sa), (cb, sb) = sims
    eng = ca.engine
    print("bricks active", eng.get_scalar(S.ACTIVE_BRICKS), "max window", eng.get_scalar(S.MAX_WINDOW_SLOTS), "overflows",
          eng.get_scalar(S.WINDOW_OVERFLOWS), flush=True)

    def fetch(c, fid):
        n_ = c.particle_num[None]
        uid = c.engine.get_field(F.UID, n_)
        a = c.engine.get_field(fid, n_)
        out = np.empty_like(a)
        out[uid] = a
        return out

    def cmp(stage, names):
        for nm in names:
            fid = getattr(F, nm)
            a, b = fetch(ca, fid), fetch(cb, fid)
            bad = np.flatnonzero((a != b).reshape(len(a), -1).any(axis=1))
            msg = f"{stage:34s} {nm:28s} differing rows {bad.size:8d}"
            if bad.size:
                d = np.abs(a.astype(np.float64) - b.astype(np.float64)).max()
                msg += f"  max |diff| {d:.3e}  first uid {bad[:5].tolist()}  a {a[bad[0]]}  b {b[bad[0]]}"
            print(msg, flush=True)

    cmp("prepare", ["DENSITY", "DFSPH_ALPHA", "POSITION", "VELOCITY"])
    n = ca.particle_num[None]
    walk_counts = fetch(ca, F.NEIGHBOR_COUNT)
    print("walk neighbour counts: max", walk_counts.max(), "mean(fluid)", walk_counts[fetch(ca, F.MATERIAL) == 1].mean(), flush=True)
    for step in range(2):
        for name, fields in [("compute_non_pressure_acceleration", ["ACCELERATION"]),
                             ("update_fluid_velocity", ["VELOCITY"]),
                             ("correct_density_error", ["VELOCITY", "DENSITY_STAR", "DFSPH_KAPPA"]),
                             ("update_fluid_position", ["POSITION"])]:
            ra, rb = getattr(sa, name)(), getattr(sb, name)()
            cmp(f"step {step} {name} {ra} {rb}", fields)
        for s_ in (sa, sb):
            s_.enforce_domain_boundary_3D(s_.container.material_fluid)
            s_.container.prepare_neighborhood_search()
            s_.compute_density()
        cmp(f"step {step} sort + compute_density", ["DENSITY"])
        for s_ in (sa, sb):
            s_.compute_alpha()
        cmp(f"step {step} compute_alpha", ["DFSPH_ALPHA"])
        ra, rb = sa.correct_divergence_error(), sb.correct_divergence_error()
        cmp(f"step {step} correct_divergence_error {ra} {rb}", ["VELOCITY", "DENSITY_DERIVATIVE"])
        print("bricks active", eng.get_scalar(S.ACTIVE_BRICKS), "max window", eng.get_scalar(S.MAX_WINDOW_SLOTS), "overflows",
              eng.get_scalar(S.WINDOW_OVERFLOWS), flush=True)


if __name__ == "__main__":
    main()
