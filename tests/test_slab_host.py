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
