"""Host-side logic of the Z-slab path on CPU: slab ranges, ownership, and the torch.distributed
plumbing over the gloo backend with world_size 2 (the NCCL data path itself needs GPUs)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from helpers import ROOT, scene
from sph_project_b200.slab import SlabContext, balanced_ranges, cell_layer


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        counts = rng.integers(0, 5000, size=50)
        r = balanced_ranges(counts, world)
        assert r[0][0] == 0 and r[-1][1] == 50
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(hi - lo >= 2 for lo, hi in r)
        loads = [counts[lo:hi].sum() for lo, hi in r]
        assert max(loads) <= counts.sum() / world + 2 * counts.max()
    with pytest.raises(ValueError):
        balanced_ranges([1, 1, 1], 2)


def test_uniform_block_splits_evenly():
    counts = [0, 0] + [1000] * 40 + [0, 0]
    r = balanced_ranges(counts, 4)
    assert [counts[lo:hi].__len__() for lo, hi in r] and [sum(counts[lo:hi]) for lo, hi in r] == [10000] * 4


def test_ownership_is_a_partition():
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 2.0, size=(20000, 3)).astype(np.float32)
    nz = 50
    counts = np.bincount(cell_layer(x[:, 2], 0.04, nz), minlength=nz)
    ranges = balanced_ranges(counts, 4)
    owned = np.stack([SlabContext(r, 4, 0.04, nz, ranges).owned(x) for r in range(4)])
    assert np.all(owned.sum(0) == 1)
    for r in range(4):
        ctx = SlabContext(r, 4, 0.04, nz, ranges)
        lo, hi = max(ctx.z_lo - 1, 0), min(ctx.z_hi + 1, nz)
        assert ctx.capacity(counts) >= counts[lo:hi].sum()


def test_cell_layer_matches_device_arithmetic():
    z = np.array([0.0, 0.0399999, 0.04, 0.07999999, 0.08, 1.9999999, 2.5], dtype=np.float32)
    expect = np.clip((z / np.float32(0.04)).astype(np.int64), 0, 49)
    assert np.array_equal(cell_layer(z, 0.04, 50), expect)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch.distributed as dist
    from helpers import scene
    from sph_project_b200.slab import SlabContext, balanced_ranges, cell_layer, broadcast_bytes
    from sph_project_b200.containers.base_container import _lattice
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # the 128-byte id rank 0 would get from ncclGetUniqueId
    uid = bytes(range(128)) if rank == 0 else None
    got = broadcast_bytes(uid, 128, src=0)
    assert got == bytes(range(128))
    # every rank derives the same slab ranges from the same scene and owns a disjoint share
    pos = _lattice([0.1, 0.1, 0.1], [0.2, 0.4, 1.0], 0.02, 3)
    nz = 30
    counts = np.bincount(cell_layer(pos[:, 2], 0.04, nz), minlength=nz)
    ctx = SlabContext(rank, world, 0.04, nz, balanced_ranges(counts, world))
    mine = int(ctx.owned(pos).sum())
    out = [None] * world
    dist.all_gather_object(out, (ctx.ranges, mine))
    assert all(o[0] == out[0][0] for o in out)
    assert sum(o[1] for o in out) == pos.shape[0]
    assert abs(out[0][1] - out[1][1]) <= counts.max()
    dist.destroy_process_group()
    print("WORKER_OK", rank)
""")


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0 and out.stdout.count("WORKER_OK") == 2, out.stdout[-1500:] + out.stderr[-3000:]


def test_halo_ranges_are_contiguous_and_aligned():
    """The invariant sph_slab.cu rests on, replayed in numpy: with a z-slowest flatten and a stable
    sort, a rank's boundary layer and its neighbour's ghost layer are contiguous index ranges holding
    the same particles in the same order, so a halo refresh is a plain range copy."""
    rng = np.random.default_rng(3)
    nx, ny, nz, h = 7, 5, 12, 0.04
    n = 6000
    x = rng.uniform(0, [nx * h, ny * h, nz * h], size=(n, 3)).astype(np.float32)
    uid = np.arange(n)
    cell = np.clip((x / np.float32(h)).astype(np.int64), 0, [nx - 1, ny - 1, nz - 1])
    flat = (cell[:, 2] * ny + cell[:, 1]) * nx + cell[:, 0]
    z_cut = 6                                               # rank 0 owns layers [0, 6), rank 1 [6, 12)

    def local_order(owned_mask, ghost_mask, prev_order):
        """pre-sort local array = previously sorted owned particles, then the imported ghosts in the
        sender's pre-sort order; the sort is a stable sort by cell."""
        ids = np.concatenate([prev_order[owned_mask[prev_order]], ghost_mask])
        return ids[np.argsort(flat[ids], kind="stable")]

    own0, own1 = cell[:, 2] < z_cut, cell[:, 2] >= z_cut
    prev0, prev1 = rng.permutation(n), rng.permutation(n)   # arbitrary previous orders on both ranks
    send0 = prev0[(own0 & (cell[:, 2] == z_cut - 1))[prev0]]   # rank 0's boundary layer in ITS pre-sort order
    send1 = prev1[(own1 & (cell[:, 2] == z_cut))[prev1]]
    order0 = local_order(own0, send1, prev0)
    order1 = local_order(own1, send0, prev1)
    # rank 0: [owned ...][ghost layer z_cut]; its top boundary layer is the tail of the owned range
    lay0 = cell[order0, 2]
    assert np.all(np.diff(lay0) >= 0)
    top0 = order0[lay0 == z_cut - 1]
    ghost0 = order0[lay0 == z_cut]
    lay1 = cell[order1, 2]
    bottom1 = order1[lay1 == z_cut]
    ghost1 = order1[lay1 == z_cut - 1]
    assert np.array_equal(uid[top0], uid[ghost1])           # rank 0's send range == rank 1's ghost range
    assert np.array_equal(uid[bottom1], uid[ghost0])
    # contiguity: each of these sets is one index range of the sorted local array
    for order, layer in ((order0, z_cut - 1), (order0, z_cut), (order1, z_cut), (order1, z_cut - 1)):
        idx = np.nonzero(cell[order, 2] == layer)[0]
        assert idx.size and idx[-1] - idx[0] + 1 == idx.size


def test_balanced_ranges_properties():
    """Property test (hypothesis): for any layer histogram the slabs are contiguous, cover every layer exactly once,
    respect the minimum thickness, and no slab is heavier than the ideal share by more than its two heaviest
    boundary layers (the cut can only move by whole layers)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 8).flatmap(lambda w: st.tuples(st.just(w), st.lists(st.integers(0, 50_000), min_size=2 * w, max_size=96))))
    def check(case):
        world, counts = case
        ranges = balanced_ranges(counts, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(counts)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(hi - lo >= 2 for lo, hi in ranges)
        total = sum(counts)
        heaviest = max(counts) if counts else 0
        # the min-thickness clamp may force weight onto a slab when the histogram is concentrated in few layers;
        # otherwise each slab stays within two layers' worth of the ideal share
        if all(hi - lo > 2 for lo, hi in ranges):
            for lo, hi in ranges:
                assert sum(counts[lo:hi]) <= total / world + 2 * heaviest

    check()


# ---------------------------------------------------------------- slab-local scene generation
def _random_boxes(seed, n):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        lower = rng.uniform(-0.3, 0.5, 3)
        extent = rng.uniform(0.15, 0.9, 3)
        space = float(rng.choice([0.02, 0.025, 0.04, 0.0125 * 4]))
        yield lower, extent, space, float(rng.choice([0.03, 0.05, 0.0]))


def test_lattice_rows_equal_the_filtered_full_lattice():
    """A Z-slab rank builds the lattice rows of its own z indices only: same positions (bitwise), same row numbers as
    building everything and filtering (what every rank did before)."""
    from sph_project_b200.containers.base_container import _lattice, _lattice_axes, _lattice_rows
    for lower, extent, space, _ in _random_boxes(1, 12):
        full = _lattice(lower, extent, space, 3)
        axes = _lattice_axes(lower, extent, space, 3)
        nzv = len(axes[2])
        for kz in (np.arange(nzv), np.arange(nzv // 3, 2 * nzv // 3), np.array([0, nzv - 1]), np.array([], dtype=np.int64)):
            pos, rows = _lattice_rows(axes, kz)
            z_index = np.arange(full.shape[0]) % nzv
            expect = np.flatnonzero(np.isin(z_index, kz))
            assert np.array_equal(rows, expect)
            assert pos.dtype == np.float32 and np.array_equal(pos, full[expect])


def test_shell_rows_equal_the_filtered_full_shell():
    from sph_project_b200.containers.base_container import (_lattice, _lattice_axes, _shell_masks, _shell_rows,
                                                            _shell_z_counts)
    cases = list(_random_boxes(2, 12))
    cases.append((np.array([0.0 + 0.04] * 3), np.array([8.5 - 0.08, 8.0 - 0.08, 2.0 - 0.08]) / 4, 0.02, 0.03))   # np.float64 corners
    cases.append(([0.04, 0.04, 0.04], [0.52, 0.72, 1.12], 0.02, 0.03))                                            # python floats
    for lower, extent, space, thickness in cases:
        full = _lattice(lower, extent, space, 3)
        mask = np.zeros(full.shape[0], dtype=bool)
        for i in range(3):   # the container's _box_shell (base_container.py:830-835 upstream)
            mask |= (full[:, i] <= lower[i] + thickness) | (full[:, i] >= lower[i] + extent[i] - thickness)
        shell = full[mask]
        shell_row_of = np.cumsum(mask) - 1
        axes = _lattice_axes(lower, extent, space, 3)
        masks = _shell_masks(axes, lower, extent, thickness)
        nzv = len(axes[2])
        per_z = _shell_z_counts(masks)
        assert per_z.sum() == shell.shape[0]
        assert np.array_equal(per_z, np.bincount(np.flatnonzero(mask) % nzv, minlength=nzv))
        for kz in (np.arange(nzv), np.arange(nzv // 4, nzv // 2), np.array([0, 1, nzv - 2, nzv - 1]), np.array([], dtype=np.int64)):
            pos, rows = _shell_rows(axes, masks, kz)
            expect_full_rows = np.flatnonzero(mask & np.isin(np.arange(full.shape[0]) % nzv, kz))
            assert np.array_equal(rows, shell_row_of[expect_full_rows])
            assert pos.dtype == np.float32 and np.array_equal(pos, full[expect_full_rows])


def test_slab_container_inserts_only_its_layers(monkeypatch):
    """The container of a slab rank (engine: the CPU oracle, slab calls faked) inserts exactly the particles the unsharded
    container holds in that rank's layers, with the unsharded insertion indices as uids; the layer histogram it cuts the
    scene by equals the histogram of the unsharded particles."""
    import copy
    import types
    from helpers import oracle_library
    from sph_project_b200._native import F
    from sph_project_b200.containers import DFSPHContainer
    from sph_project_b200.utils import SimConfig
    sc = scene("dfsph", domain_end=(0.6, 0.8, 1.2), block_start=(0.1, 0.1, 0.1), block_end=(0.3, 0.5, 1.1), dt=1e-3)
    lib = oracle_library()

    def build():
        return DFSPHContainer(SimConfig(config=copy.deepcopy(sc), verbose=False), GGUI=False, engine_library=lib)

    ref = build()
    ref.insert_object()
    n = ref.particle_num[None]
    x_ref, obj_ref = ref.engine.get_field(F.POSITION, n), ref.engine.get_field(F.OBJECT_ID, n)
    assert np.array_equal(ref.engine.get_field(F.UID, n), np.arange(n))
    nz = int(ref.grid_num[2])
    layers = cell_layer(x_ref[:, 2], ref.dh, nz)

    hist = sum(c for c, _ in ref._scene_layers(nz))
    assert np.array_equal(hist, np.bincount(layers, minlength=nz))
    fluid_hist = sum(c for c, is_fluid in ref._scene_layers(nz) if is_fluid)
    assert fluid_hist.sum() == ref.fluid_particle_num[None]

    # the slab calls of the library, faked on the oracle's engine class (the oracle always holds the whole domain)
    import sph_project_b200._native as nat
    import sph_project_b200.slab as slab_mod
    from sph_project_b200.containers.base_container import BaseContainer
    totals = []
    monkeypatch.setattr(nat.Engine, "slab_unique_id", lambda self: bytes(128))
    monkeypatch.setattr(nat.Engine, "slab_init", lambda self, rank, world, uid, z_lo, z_hi, n_global: setattr(self, "_fake_range", (z_lo, z_hi)))
    monkeypatch.setattr(nat.Engine, "slab_info", lambda self: types.SimpleNamespace(z_lo=self._fake_range[0], z_hi=self._fake_range[1]))
    monkeypatch.setattr(nat.Engine, "slab_set_global_particle_num", lambda self, v: totals.append(v))
    monkeypatch.setattr(slab_mod, "broadcast_bytes", lambda payload, nbytes, src=0: payload)
    monkeypatch.setattr(BaseContainer, "_connect_slab_peers", lambda self: None)
    make_params = BaseContainer._make_params
    monkeypatch.setattr(BaseContainer, "_make_params", lambda self, device, slab: make_params(self, device, False))

    world, seen, cuts = 3, 0, None
    for rank in range(world):
        c = DFSPHContainer(SimConfig(config=copy.deepcopy(sc), verbose=False), GGUI=False, engine_library=lib, slab=(rank, world))
        assert cuts is None or cuts == c.slab.ranges
        cuts = c.slab.ranges
        assert np.array_equal(c._layer_counts, hist)
        z_lo, z_hi = cuts[rank]
        c.insert_object()
        m = c.particle_num[None]
        mine = np.flatnonzero((layers >= z_lo) & (layers < z_hi))
        assert m == mine.size and totals[-1] == n
        assert np.array_equal(c.engine.get_field(F.UID, m), mine)
        assert np.array_equal(c.engine.get_field(F.POSITION, m), x_ref[mine])
        assert np.array_equal(c.engine.get_field(F.OBJECT_ID, m), obj_ref[mine])
        assert c.fluid_particle_num[None] == ref.fluid_particle_num[None]   # scene totals, as before
        seen += m
    assert seen == n and cuts[0][0] == 0 and cuts[-1][1] == nz


def test_box_particle_count_without_building_the_box():
    from helpers import oracle_library
    from sph_project_b200.containers import DFSPHContainer
    from sph_project_b200.utils import SimConfig
    c = DFSPHContainer(SimConfig(config=scene("dfsph", domain_end=(0.7, 0.5, 0.9), dt=1e-3), verbose=False), GGUI=False,
                       engine_library=oracle_library())
    for lower, size, t, space in (([0.04] * 3, [0.62, 0.42, 0.82], 0.03, 0.02), (c.domain_box_start, c.domain_box_size, 0.03, 0.02),
                                  ([0.0, 0.1, 0.2], [0.3, 0.3, 0.3], 0.05, 0.025)):
        assert c.compute_box_particle_num(lower, size, t, space) == c._box_shell(lower, size, t, space).shape[0]
