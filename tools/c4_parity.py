#!/usr/bin/env python
"""C4 (PCISPH + implicit viscosity) at full size: brick lists vs global walk vs CPU oracle from the same state."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    sc = bench.scene_for("c4_buckling")
    cg, sg = bench.make_sim(sc)
    cg.engine.step(50)
    xs, vs, mats = bench.fields_by_uid(cg, with_material=True)
    os.environ["SPH_B200_NO_LISTS"] = "1"
    cw, sw = bench.make_sim(sc)
    os.environ.pop("SPH_B200_NO_LISTS")
    co, so = bench.make_sim(sc, bench.oracle_library())
    for c, s in ((cg, sg), (cw, sw), (co, so)):
        bench.load_state(c, s, xs, vs, mats)
    fl = None
    for k in range(steps):
        st = [s.step(1) for s in (sg, sw, so)]
        x = [bench.fields_by_uid(c)[0] for c in (cg, cw, co)]
        if fl is None:
            from sph_project_b200._native import F
            n = cg.particle_num[None]
            uid = cg.engine.get_field(F.UID, n)
            mat = np.empty(n, np.int32); mat[uid] = cg.engine.get_field(F.MATERIAL, n)
            fl = mat == 1
        scale = np.abs(x[2]).max()
        d_gw = np.abs(x[0] - x[1]).max() / scale
        d_go = np.abs(x[0] - x[2]).max() / scale
        disp = np.abs(x[2][fl] - xs[fl]).max()
        print(f"step {k + 1}: cg iters lists/walk/oracle {st[0].cg_iterations}/{st[1].cg_iterations}/{st[2].cg_iterations}  cg_err {st[0].cg_error:.2e}/{st[2].cg_error:.2e}  "
              f"pcisph {st[0].pcisph_iterations}/{st[2].pcisph_iterations}  rel dx lists-walk {d_gw:.2e}  lists-oracle {d_go:.2e}  (max displacement so far {disp:.3e} m, scale {scale:.1f} m)", flush=True)


if __name__ == "__main__":
    main()
