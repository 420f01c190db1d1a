"""Host-side plumbing of the Z-slab decomposition; no reference counterpart.

One process per GPU (torchrun).  Every rank builds the same scene description, keeps only the
particles whose cell layer falls into its slab, and the library does the rest over NCCL
(migration, ghost import, halo refreshes, error all-reduces: sph_project_b200/csrc/sph_slab.cu).
The host side is small on purpose:

  * `balanced_ranges`   split the grid's cell layers into `world` contiguous slabs with about the
                        same number of particles each (from a cell-layer histogram);
  * `broadcast_bytes`   ship the 128-byte NCCL unique id from rank 0 with torch.distributed
                        (works on the `nccl` and on the `gloo` backend);
  * `SlabContext`       rank / world / ranges / ownership test used by the particle container.
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

    def owned(self, positions: np.ndarray) -> np.ndarray:
        cz = cell_layer(positions[:, 2], self.dh, self.nz)
        return (cz >= self.z_lo) & (cz < self.z_hi)

    def capacity(self, layer_counts: Sequence[int], slack: float = 1.3, extra: int = 65536) -> int:
        """Local particle capacity: owned layers + one ghost layer each side, with head-room for
        migration imbalance."""
        c = np.asarray(layer_counts, dtype=np.int64)
        lo, hi = max(self.z_lo - 1, 0), min(self.z_hi + 1, self.nz)
        return int(c[lo:hi].sum() * slack) + extra
