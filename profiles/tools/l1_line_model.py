"""L1 line-lookup model of the list-consumer kernels (a MODEL computed on the host, not a measurement).

ncu shows the two DFSPH iteration kernels at 80-91 % of L1 throughput with ~1 sector per accepted pair
(profiles/r01_ncu_v4_record_gathers_raw.csv); bench.py's `l1_gather` figure puts them at 0.83-0.93 pairs per
SM-clock against a bound of one 128-byte line lookup per clock per SM.  This script asks what the access
pattern itself allows: for the benchmark's particle layout (1.23 M-fluid dam break, the CUDA path's
x-fastest cell order, in-cell insertion order, 32-byte records = 4 per line) it counts the distinct
128-byte lines each warp-wide gather touches under

  thread-per-particle  lane l of a warp handles particle 32w+l, request k gathers nbr[k][32w+l]   (what ships)
  W lanes per particle W consecutive entries of ONE particle's (ascending) list per request, W = 32, 16, 8

Neighbour lists come from the CPU oracle (same neighbour sets as the CUDA path, tests/test_gpu_parity.py).

    python profiles/tools/l1_line_model.py [steps] [jitter]
        steps   the oracle runs first (default 3: the near-lattice early window)
        jitter  > 0: displace every fluid particle by uniform +-jitter x diameter and shuffle the in-cell order,
                a stand-in for the disordered pressurised window (1000 oracle steps are out of reach on the host)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

REC_PER_LINE = 4      # 32-byte records in a 128-byte line


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    c, s = bench.make_sim(bench.dam_break_scene("dfsph"), bench.oracle_library())
    jitter = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    if steps:
        s.step(steps)
    n = int(c.particle_num[None])
    rng = np.random.default_rng(0)
    if jitter > 0:
        x = c.particle_positions.to_numpy(n)
        m = c.particle_materials.to_numpy(n) == 1
        x[m] += rng.uniform(-jitter, jitter, size=(int(m.sum()), 3)).astype(np.float32) * np.float32(c.particle_diameter)
        c.particle_positions.from_numpy(x)
    c.prepare_neighborhood_search()
    x = c.particle_positions.to_numpy(n)
    mat = c.particle_materials.to_numpy(n)
    uid = c.particle_uids.to_numpy(n)
    if jitter > 0:
        uid = rng.permutation(n)
    off, idx = c.engine.get_neighbors()
    # the CUDA path's order: x-fastest flatten, insertion order inside a cell
    g = np.asarray(c.grid_num)
    cell = np.minimum(np.maximum((x / np.float32(c.grid_size)).astype(np.int64), 0), g - 1)
    flat = (cell[:, 2] * g[1] + cell[:, 1]) * g[0] + cell[:, 0]
    order = np.lexsort((uid, flat))
    new_of_old = np.empty(n, dtype=np.int64)
    new_of_old[order] = np.arange(n)
    cnt_old = np.diff(off).astype(np.int64)
    cnt = cnt_old[order]
    fluid = (mat == 1)[order]
    kmax = int(cnt.max())
    # ELL in the new order, each list ascending in the new index
    ell = np.full((n, kmax), -1, dtype=np.int64)
    rows_old = np.repeat(np.arange(n), cnt_old)
    col = np.arange(idx.size) - np.repeat(off[:-1].astype(np.int64), cnt_old)
    ell[new_of_old[rows_old], col] = new_of_old[idx]
    ell.sort(axis=1)                                   # -1 padding first, then ascending j
    ell = np.where(ell >= 0, ell, np.iinfo(np.int64).max)
    ell.sort(axis=1)
    ell[ell == np.iinfo(np.int64).max] = -1
    ell[~fluid] = -1                                   # non-fluid rows do no gathers in the iteration kernels
    pairs = int((ell >= 0).sum())
    print(f"particles {n}, fluid rows {int(fluid.sum())}, accepted pairs {pairs}, mean list {pairs / fluid.sum():.1f}, max {kmax}")

    def distinct_per_group(lines):
        """lines: (groups, width) int64 with -1 = idle lane -> total distinct valid lines over groups."""
        srt = np.sort(lines, axis=1)
        valid = srt >= 0
        new = valid & np.concatenate([np.ones((srt.shape[0], 1), bool), srt[:, 1:] != srt[:, :-1]], axis=1)
        return int(new.sum())

    pad = (-n) % 32
    lines_tpp = 0
    requests_tpp = 0
    for k in range(kmax):
        colk = np.concatenate([ell[:, k], np.full(pad, -1, dtype=np.int64)])
        ln = np.where(colk >= 0, colk // REC_PER_LINE, -1).reshape(-1, 32)
        lines_tpp += distinct_per_group(ln)
        requests_tpp += int((ln >= 0).any(axis=1).sum())
    print(f"thread-per-particle : {lines_tpp / pairs:.3f} lines per pair, {requests_tpp} warp requests, "
          f"{pairs / requests_tpp:.1f} active lanes per request")
    for w in (32, 16, 8, 4):
        padk = (-kmax) % w
        e = np.concatenate([ell, np.full((n, padk), -1, dtype=np.int64)], axis=1)[fluid]
        ln = np.where(e >= 0, e // REC_PER_LINE, -1).reshape(-1, w)
        lines = distinct_per_group(ln)
        req = int((ln >= 0).any(axis=1).sum()) * w // 32      # warp requests when 32 / w particles share a warp
        print(f"{w:2d} lanes per particle: {lines / pairs:.3f} lines per pair, ~{req} warp requests, "
              f"{pairs / max(req, 1):.1f} active lanes per request")


if __name__ == "__main__":
    main()
