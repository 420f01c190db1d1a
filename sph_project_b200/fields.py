"""Stand-ins for the Taichi fields the reference exposes on its containers and solvers.

The reference's callers touch fields as `field.to_numpy()`, `field[None]`, `field[i]`,
`field[i] = v`, `field.fill(v)` and `field.from_numpy(a)` (e.g. base_container.py:599-609,
bullet_solver.py:144-167, run_simulation.py:139-144).  These classes give the same access on
top of the C ABI; the data itself lives in device memory owned by the library.
"""
from __future__ import annotations

import numpy as np

from ._native import FIELD_LAYOUT, MAX_OBJECTS


class ParticleField:
    """Per-particle field of length particle_max_num (ti.field / ti.Vector.field / ti.Matrix.field)."""

    def __init__(self, engine, field_id: int, capacity: int, matrix: bool = False):
        self._engine = engine
        self.field_id = field_id
        self.capacity = capacity
        self.components, self.dtype = FIELD_LAYOUT[field_id]
        self._matrix = matrix
        if self.components == 1:
            self.shape = (capacity,)
        else:
            self.shape = (capacity,)  # Taichi reports the field shape without the vector dimension

    def to_numpy(self, n: int | None = None) -> np.ndarray:
        """Whole field (all particle_max_num slots, like Taichi) or its first n entries."""
        a = self._engine.get_field(self.field_id, self.capacity if n is None else n)
        if self._matrix:
            a = a.reshape(-1, 3, 3)
        return a

    def from_numpy(self, values: np.ndarray):
        self._engine.set_field(self.field_id, np.asarray(values).reshape(-1))

    def fill(self, value):
        self._engine.fill_field(self.field_id, value)

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            a = self._engine.get_field(self.field_id, int(index) + 1)
            v = a[int(index)]
            return v.reshape(3, 3) if self._matrix else v
        return self.to_numpy()[index]

    def __setitem__(self, index, value):
        n = self.capacity
        a = self._engine.get_field(self.field_id, n)
        a[index] = np.asarray(value, dtype=self.dtype).reshape(a[index].shape)
        self._engine.set_field(self.field_id, a)

    def __len__(self):
        return self.capacity


class ScalarField:
    """0-d field backed by a library scalar: `field[None]` reads / writes it."""

    def __init__(self, engine, scalar_id: int, cast=float):
        self._engine = engine
        self.scalar_id = scalar_id
        self._cast = cast
        self.shape = ()

    def __getitem__(self, index):
        assert index is None, "0-d field: index with [None]"
        return self._cast(self._engine.get_scalar(self.scalar_id))

    def __setitem__(self, index, value):
        assert index is None, "0-d field: index with [None]"
        self._engine.set_scalar(self.scalar_id, value)

    def to_numpy(self):
        return np.asarray(self[None])


class HostScalar:
    """0-d field that only the Python host uses (e.g. object_num)."""

    def __init__(self, value=0):
        self._value = value
        self.shape = ()

    def __getitem__(self, index):
        return self._value

    def __setitem__(self, index, value):
        self._value = value

    def to_numpy(self):
        return np.array(self._value)


class HostField:
    """A field that lives on the host only (the GGUI staging buffers): numpy storage behind the field idioms."""

    def __init__(self, shape, dtype=np.float32):
        self._a = np.zeros(shape, dtype=dtype)
        self.shape = (self._a.shape[0],)

    def to_numpy(self):
        return self._a.copy()

    def from_numpy(self, a):
        self._a[...] = a

    def fill(self, value):
        self._a[...] = value

    def __getitem__(self, index):
        return self._a[index]

    def __setitem__(self, index, value):
        self._a[index] = value


class ObjectTable:
    """Per-object table (shape max_num_object) kept on the host and mirrored into the library
    through `push(object_id)` whenever an entry changes (object_materials, rigid_body_*)."""

    def __init__(self, shape_tail=(), dtype=np.float32, push=None):
        self._a = np.zeros((MAX_OBJECTS,) + tuple(shape_tail), dtype=dtype)
        self._push = push
        self.shape = (MAX_OBJECTS,)

    def __getitem__(self, index):
        v = self._a[index]
        return v.item() if np.ndim(v) == 0 else v

    def __setitem__(self, index, value):
        self._a[index] = np.asarray(value, dtype=self._a.dtype).reshape(self._a[index].shape)
        if self._push is not None:
            self._push(int(index))

    def to_numpy(self):
        return self._a.copy()

    def fill(self, value):
        self._a[...] = value
        if self._push is not None:
            for i in range(MAX_OBJECTS):
                self._push(i)


class _WrenchState:
    """What the host has taken off the device accumulators.  The library can only zero every object's wrench
    at once, the reference zeroes one object and one table at a time while it walks its bodies
    (bullet_solver.py:149-156): a per-object reset moves the device sums into this host-side remainder and clears
    the one entry there."""

    def __init__(self, engine):
        self.engine = engine
        self.remainder = None          # [2][MAX_OBJECTS][3] float32, or None while nothing is held back

    def read(self):
        dev = np.stack(self.engine.get_rigid_wrench())
        return dev if self.remainder is None else dev + self.remainder

    def reset(self, which, index):
        total = self.read()
        self.engine.zero_rigid_wrench()
        total[which][index] = 0.0
        self.remainder = total if np.any(total != 0) else None

    def reset_all(self, which):
        total = self.read()
        self.engine.zero_rigid_wrench()
        total[which] = 0.0
        self.remainder = total if np.any(total != 0) else None


class WrenchTable:
    """rigid_body_forces / rigid_body_torques: accumulated on the device by the fluid kernels, read and reset here.
    The two tables of a container share one `_WrenchState`."""

    def __init__(self, engine, which: int, state: "_WrenchState" = None):
        self._state = state if state is not None else _WrenchState(engine)
        self._which = which
        self.shape = (MAX_OBJECTS,)

    @property
    def state(self):
        return self._state

    def to_numpy(self):
        return self._state.read()[self._which]

    def __getitem__(self, index):
        return self.to_numpy()[index]

    def __setitem__(self, index, value):
        # the reference only ever writes zeros here (bullet_solver.py:155-156)
        if np.any(np.asarray(value) != 0):
            raise NotImplementedError("rigid wrench accumulators can only be reset")
        self._state.reset(self._which, index)

    def fill(self, value):
        if np.any(np.asarray(value) != 0):
            raise NotImplementedError("rigid wrench accumulators can only be reset")
        self._state.reset_all(self._which)
