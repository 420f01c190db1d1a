"""The host-side replay of the brick layout (profiles/tools/smem_bank_model.py: cell sort, bricks, window slot numbering,
slot lists in walk order) decodes to exactly the oracle's neighbour sets: the model under profiles/ counts the accesses of
the real layout, not of something like it."""
import os
import sys

import numpy as np

from helpers import ROOT, make_sim, oracle_library, scene

sys.path.insert(0, os.path.join(ROOT, "profiles", "tools"))


def test_brick_slot_lists_decode_to_the_oracle_neighbour_sets():
    import smem_bank_model as model
    sc = scene("dfsph", domain_end=(0.7, 0.9, 0.5), block_start=(0.1, 0.1, 0.1), block_end=(0.4, 0.6, 0.4), dt=1e-3,
               velocity=(0.3, -1.0, 0.2))
    c, s = make_sim(sc, oracle_library())
    s.step(25)
    c.prepare_neighborhood_search()
    n = int(c.particle_num[None])
    x = c.particle_positions.to_numpy(n)
    fluid = c.particle_materials.to_numpy(n) == 1
    uid = c.particle_uids.to_numpy(n)
    off, idx = c.engine.get_neighbors()
    order, bricks = model.brick_lists(x, fluid, uid, np.asarray(c.grid_num, dtype=np.int64), np.float32(c.grid_size), float(c.dh))
    seen = np.zeros(n, dtype=bool)
    for rows, lists, slot_to_index, window in bricks:
        assert window == slot_to_index.size and window < 65536
        for i, slots in zip(rows, lists):
            old_i = order[i]
            assert fluid[old_i] and not seen[old_i]
            seen[old_i] = True
            got = np.sort(order[slot_to_index[slots]])
            assert np.array_equal(got, np.sort(idx[off[old_i]:off[old_i + 1]]))
    assert np.array_equal(seen, fluid)     # every fluid particle is a row of exactly one brick
