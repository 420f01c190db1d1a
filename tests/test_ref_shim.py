"""The Taichi emulation behind tests/golden/ref_*.npz (tests/golden/ref_shim/taichi) keeps the semantics its
docstring claims -- the fixtures are only as good as this."""
import importlib.util
import os

import numpy as np

from helpers import ROOT

_spec = importlib.util.spec_from_file_location("_ref_shim_taichi", os.path.join(ROOT, "tests", "golden", "ref_shim", "taichi", "__init__.py"))
ti = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ti)


@ti.data_oriented
class Toy:
    def __init__(self, n=8):
        self.n = n
        self.x = ti.Vector.field(3, dtype=float, shape=n)
        self.s = ti.field(dtype=float, shape=n)
        self.cnt = ti.field(int, shape=4)
        self.out = ti.field(dtype=float, shape=())
        self.scale = 0.1            # a Python float captured from kernel scope -> f32
        self.dim = 3

    @ti.func
    def add_to(self, i, ret: ti.template()):
        ret += self.s[i]

    @ti.func
    def bump_vec(self, ret: ti.template()):
        ret += ti.Vector([1.0, 2.0, 3.0])

    @ti.kernel
    def by_reference(self) -> ti.f32:
        acc = 0.0
        for i in range(self.n):
            self.add_to(i, acc)
        v = ti.Vector([0.0 for _ in range(self.dim)])
        self.bump_vec(v)
        self.bump_vec(v)
        return acc + v[2]

    @ti.kernel
    def f32_rounding(self) -> ti.f32:
        a = self.scale
        b = a * 3.0
        return b

    @ti.kernel
    def copies_and_views(self) -> ti.f32:
        p = self.x[1]                 # local copy
        self.x[1][0] = 5.0            # write-through element store
        self.x[1] += ti.Vector([0.0, 1.0, 0.0])
        q = p
        q[2] = 9.0                    # q is its own copy, p unchanged
        return p[0] * 100.0 + p[2] * 10.0 + self.x[1][0] + self.x[1][1]

    @ti.kernel
    def atomics(self) -> int:
        self.cnt.fill(0)
        for i in range(6):
            ti.atomic_add(self.cnt[i % 2], 1)
        old = ti.atomic_sub(self.cnt[0], 1)
        n = 0
        n += old                       # an int variable stays an int
        for I in ti.grouped(self.cnt):
            self.cnt[I] += 10
        return n

    @ti.kernel
    def struct_and_ranges(self) -> ti.f32:
        ret = ti.Struct(total=0.0, hits=0)
        for off in ti.grouped(ti.ndrange(*((-1, 2),) * 2)):
            ret.total += off[0] * 10.0 + off[1]
            ret.hits += 1
        first = 0.0
        k = 0
        for off in ti.grouped(ti.ndrange((-1, 2), (-1, 2))):
            if k == 1:
                first = off[0] * 10.0 + off[1]     # second tuple: last index fastest -> (-1, 0)
            k += 1
        return ret.total + ret.hits * 1000.0 + first * 100000.0


def test_template_arguments_are_by_reference():
    t = Toy()
    t.s.from_numpy(np.arange(8, dtype=np.float32))
    assert t.by_reference() == 28.0 + 6.0


def test_float_constants_round_to_f32():
    got = Toy().f32_rounding()
    assert got == float(np.float32(0.1) * np.float32(3.0))
    assert got != 0.1 * 3.0


def test_locals_copy_and_field_elements_write_through():
    t = Toy()
    t.x.from_numpy(np.ones((8, 3), dtype=np.float32))
    assert t.copies_and_views() == 100.0 + 10.0 + 5.0 + 2.0
    assert np.array_equal(t.x.to_numpy()[1], np.array([5.0, 2.0, 1.0], dtype=np.float32))


def test_atomics_return_the_old_value_and_ints_stay_ints():
    t = Toy()
    n = t.atomics()
    assert n == 3 and isinstance(n, int)
    assert t.cnt.to_numpy().tolist() == [12, 13, 10, 10]


def test_struct_and_ndrange_order():
    assert Toy().struct_and_ranges() == 0.0 + 9 * 1000.0 + (-10.0) * 100000.0


def test_prefix_sum_is_inclusive():
    f = ti.field(int, shape=5)
    f.from_numpy(np.array([1, 0, 2, 0, 3], dtype=np.int32))
    ti.algorithms.PrefixSumExecutor(5).run(f)
    assert f.to_numpy().tolist() == [1, 1, 3, 3, 6]


def test_out_of_bounds_reads_raise():
    import pytest
    f = ti.field(int, shape=3)
    with pytest.raises(IndexError):
        f[3]
    with pytest.raises(IndexError):
        f[-1]
