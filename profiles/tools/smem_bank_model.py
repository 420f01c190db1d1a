"""Shared-memory wavefront model of the brick sweeps (a MODEL computed on the host, not a measurement).

ncu on B200 (profiles/r02_ncu_brick_iter_raw.csv) shows the two DFSPH iteration sweeps spending ~100 us per launch in the
shared-memory pipe: 8.0-8.2 wavefronts per LDS.128 and 6.6 per LDS.64 where 4 and 2 would be conflict-free.  This script
replays the brick layout of sph_brick.cuh on a settled DFSPH state from the CPU oracle -- x-fastest cell sort, 4x4x3-cell
bricks, the 30 runs of a brick's window numbered into slots, 16-bit slot lists in walk order, working rows compacted in
owned order, 32 consecutive rows per warp -- and counts, for every warp-wide shared-memory load, the wavefronts the
access pattern itself needs:

    a 16-byte load is served in 4 phases of 8 lanes, an 8-byte load in 2 phases of 16 lanes; within a phase, lanes
    reading the same slot share a wavefront, lanes reading different slots with equal (slot mod 8) -- the same four banks --
    need one wavefront each.

It does so for what ships (one thread per row, lane l walks the list of row l) and for the layouts a next round could
try now that the lists are row-major (W lanes share one row and read W consecutive list entries, W = 32, 16, 8).

    python profiles/tools/smem_bank_model.py [settle_steps]      (default 600; ~2 min of CPU)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

BX, BY, BZ = 4, 4, 3          # SPH_BRICK_X / Y / Z
RY, RZ = BY + 2, BZ + 2       # runs per brick: (BY + 2) x (BZ + 2), each BX + 2 cells long


def settled_state(steps):
    cfg = {"domainStart": [0.0, 0.0, 0.0], "domainEnd": [1.0, 1.2, 0.6], "particleRadius": 0.01, "addDomainBox": True,
           "density0": 1000, "gravitation": [0.0, -9.81, 0.0], "simulationMethod": "dfsph", "viscosityMethod": "standard",
           "timeStepSize": 6e-4, "viscosity": 10.0, "viscosity_b": 0.3, "exportFrame": False, "exportPly": False, "exportObj": False}
    block = {"objectId": 0, "start": [0.09, 0.1, 0.1], "end": [0.49, 1.1, 0.5], "translation": [0.0, 0.0, 0.0], "scale": [1, 1, 1],
             "velocity": [0.0, -0.5, 0.0], "density": 1000.0, "color": [50, 100, 200], "entryTime": -1.0}
    c, s = bench.make_sim({"Configuration": cfg, "FluidBlocks": [block]}, bench.oracle_library())
    its = 0
    for _ in range(steps):
        st = c.engine.step(1)
        its = st.dfsph_density_iterations if hasattr(st, "dfsph_density_iterations") else its
    c.prepare_neighborhood_search()
    n = int(c.particle_num[None])
    return (c.particle_positions.to_numpy(n), c.particle_materials.to_numpy(n) == 1, c.particle_uids.to_numpy(n),
            np.asarray(c.grid_num, dtype=np.int64), np.float32(c.grid_size), float(c.dh) if hasattr(c, "dh") else float(c.grid_size), its)


def wavefronts(slots, lanes_per_phase):
    """slots: (instructions, 32) int64, -1 = inactive lane.  Wavefronts per instruction under the phase model."""
    ins = slots.shape[0]
    total = np.zeros(ins, dtype=np.int64)
    for p in range(32 // lanes_per_phase):
        s = np.sort(slots[:, p * lanes_per_phase:(p + 1) * lanes_per_phase], axis=1)
        distinct = (s >= 0) & np.concatenate([np.ones((ins, 1), bool), s[:, 1:] != s[:, :-1]], axis=1)
        worst = np.zeros(ins, dtype=np.int64)
        for g in range(8):
            worst = np.maximum(worst, (distinct & (s % 8 == g)).sum(axis=1))
        total += worst
    return total


def brick_lists(x, fluid, uid, g, cell_size, h):
    """Replay of the brick layout on a particle state.  Returns (order, bricks): `order` = the sort permutation, `bricks`
    = per active brick (rows, lists, slot_to_index, window_slots) with rows = sorted indices of its fluid particles in
    owned order, lists[k] = window slots of row k's neighbours in walk order, slot_to_index = sorted index of each slot."""
    n = x.shape[0]
    cell = np.minimum(np.maximum((x / cell_size).astype(np.int64), 0), g - 1)
    flat = (cell[:, 2] * g[1] + cell[:, 1]) * g[0] + cell[:, 0]
    order = np.lexsort((uid, flat))
    x, fluid, cell, flat = x[order], fluid[order], cell[order], flat[order]
    ncell = int(g.prod())
    start = np.searchsorted(flat, np.arange(ncell + 1))          # cell_start of the sorted arrays
    h2 = np.float32(h) * np.float32(h)
    bricks = []
    nb = [-(-int(g[0]) // BX), -(-int(g[1]) // BY), -(-int(g[2]) // BZ)]
    for bz in range(nb[2]):
        for by in range(nb[1]):
            for bx in range(nb[0]):
                x0, y0, z0 = bx * BX, by * BY, bz * BZ
                # runs: r = rz * RY + ry covers cells (x0 - 1 .. x0 + BX, y0 - 1 + ry, z0 - 1 + rz), clipped to the grid
                run_lo, run_hi = np.zeros(RY * RZ, np.int64), np.zeros(RY * RZ, np.int64)
                for rz in range(RZ):
                    for ry in range(RY):
                        cy, cz = y0 - 1 + ry, z0 - 1 + rz
                        if 0 <= cy < g[1] and 0 <= cz < g[2]:
                            xa, xb = max(x0 - 1, 0), min(x0 + BX, int(g[0]) - 1)
                            base = (cz * g[1] + cy) * g[0]
                            run_lo[rz * RY + ry], run_hi[rz * RY + ry] = start[base + xa], start[base + xb + 1]
                S = np.concatenate([[0], np.cumsum(run_hi - run_lo)])
                own = []
                for rz in range(1, RZ - 1):
                    for ry in range(1, RY - 1):
                        cy, cz = y0 - 1 + ry, z0 - 1 + rz
                        if cy >= g[1] or cz >= g[2]:
                            continue
                        base = (cz * g[1] + cy) * g[0]
                        a, b = start[base + min(x0, int(g[0]))], start[base + min(x0 + BX, int(g[0]))]
                        own.append(np.arange(a, b))
                own = np.concatenate(own) if own else np.zeros(0, np.int64)
                own = own[fluid[own]]
                if own.size == 0:
                    continue
                brick_rows = []
                for i in own:
                    cx, cy, cz = cell[i]
                    ly, lz = cy - y0, cz - z0
                    out = []
                    for k in range(9):
                        r = (lz + k // 3) * RY + (ly + k % 3)
                        yy, zz = y0 - 1 + ly + k % 3, z0 - 1 + lz + k // 3
                        if not (0 <= yy < g[1] and 0 <= zz < g[2]):
                            continue
                        base = (zz * g[1] + yy) * g[0]
                        a, b = start[base + max(cx - 1, 0)], start[base + min(cx + 1, int(g[0]) - 1) + 1]
                        j = np.arange(a, b)
                        d = x[j] - x[i]
                        acc = ((d * d).sum(axis=1) < h2) & (j != i)
                        out.append(S[r] + (j[acc] - run_lo[r]))
                    brick_rows.append(np.concatenate(out) if out else np.zeros(0, np.int64))
                slot_to_index = np.concatenate([np.arange(run_lo[r], run_hi[r]) for r in range(RY * RZ)])
                bricks.append((own, brick_rows, slot_to_index, int(S[-1])))
    return order, bricks


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    x, fluid, uid, g, cell_size, h, _ = settled_state(steps)
    n = x.shape[0]
    _, bricks = brick_lists(x, fluid, uid, g, cell_size, h)
    lists = [b[1] for b in bricks]
    rows_per_brick = [len(b[1]) for b in bricks]
    window_sizes = [b[3] for b in bricks]

    counts = np.array([len(r) for b in lists for r in b])
    pairs = int(counts.sum())
    print(f"state after {steps} oracle steps: {n} particles, {int(fluid.sum())} fluid rows in {len(lists)} bricks "
          f"({np.mean(rows_per_brick):.0f} rows per brick, windows {np.mean(window_sizes):.0f} slots on average, max {max(window_sizes)}), "
          f"{pairs} pairs, {counts.mean():.1f} neighbours per row (max {counts.max()})")

    def layout(width):
        """width lanes share a row (32 // width rows per warp-wide load); returns the (instructions, 32) slot table."""
        table = []
        per = 32 // width
        for b in lists:
            for w0 in range(0, len(b), per if width < 32 else 1):
                group = b[w0:w0 + per]
                if width == 1:
                    group = b[w0:w0 + 32]
                depth = max(len(r) for r in group)
                if width == 1:
                    for k in range(depth):
                        table.append([r[k] if k < len(r) else -1 for r in group] + [-1] * (32 - len(group)))
                else:
                    for k0 in range(0, depth, width):
                        line = []
                        for r in group:
                            seg = list(r[k0:k0 + width])
                            line += seg + [-1] * (width - len(seg))
                        table.append(line + [-1] * (32 - len(line)))
            # (thread-per-row: w0 advances by 32 rows)
        return np.array(table, dtype=np.int64)

    # thread-per-row needs its own stride over rows
    def thread_per_row():
        table = []
        for b in lists:
            for w0 in range(0, len(b), 32):
                group = b[w0:w0 + 32]
                for k in range(max(len(r) for r in group)):
                    table.append([r[k] if k < len(r) else -1 for r in group] + [-1] * (32 - len(group)))
        return np.array(table, dtype=np.int64)

    def bank_ordered():
        """Thread per row as shipped, but the list build orders each row's entries so that lane l reads bank group
        (l + k) mod 8 at step k whenever the row still has an entry of that group (leftovers fill the gaps in walk order):
        the 8 lanes of a phase then touch 8 different groups.  Only the build changes; the sums run in another order."""
        table = []
        for b in lists:
            for w0 in range(0, len(b), 32):
                group = []
                for lane, r in enumerate(b[w0:w0 + 32]):
                    queues = [[s_ for s_ in r if s_ % 8 == q] for q in range(8)]
                    out, holes = [], []
                    for k in range(len(r)):
                        q = queues[(lane + k) % 8]
                        if q:
                            out.append(q.pop(0))
                        else:
                            out.append(-1)
                            holes.append(k)
                    rest = [s_ for q in queues for s_ in q]
                    for k, s_ in zip(holes, rest):
                        out[k] = s_
                    group.append(out)
                for k in range(max(len(r) for r in group)):
                    table.append([r[k] if k < len(r) else -1 for r in group] + [-1] * (32 - len(group)))
        return np.array(table, dtype=np.int64)

    print(f"{'layout':34s} {'loads':>9s} {'lanes busy':>10s} {'wavefronts per LDS.128':>23s} {'per LDS.64':>11s} {'wavefronts per pair (one of each)':>34s}")
    for name, table in (("thread per row (ships)", thread_per_row()), ("32 lanes per row", layout(32)), ("16 lanes per row", layout(16)),
                        ("8 lanes per row", layout(8)), ("thread per row, bank-ordered lists", bank_ordered())):
        w128, w64 = wavefronts(table, 8), wavefronts(table, 16)
        busy = (table >= 0).sum() / table.size
        print(f"{name:34s} {table.shape[0]:9d} {busy:10.2f} {w128.mean():23.2f} {w64.mean():11.2f} {(w128.sum() + w64.sum()) / pairs:34.3f}")
    print("conflict-free: 4 per LDS.128, 2 per LDS.64, (4 + 2) / 32 = 0.188 per pair with every lane busy")


if __name__ == "__main__":
    main()
