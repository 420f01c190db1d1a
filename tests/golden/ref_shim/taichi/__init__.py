"""Just enough of the Taichi API to EXECUTE the reference's Python sources in this container.

Test infrastructure, used only by tests/golden/make_ref_golden.py.  Taichi itself cannot be installed
offline, so the reference's `@ti.kernel` / `@ti.func` bodies are re-parsed (ast) and run as plain serial
Python with Taichi's value semantics restated on numpy scalars:

  * default_fp = f32, default_ip = i32 (run_simulation.py:9 calls ti.init without overrides): every float
    that enters kernel scope -- literal, Python/numpy constant captured from `self`, field element --
    becomes np.float32 and every operation rounds to f32 (numpy >= 2 scalar promotion);
  * locals are value copies (`pos = field[i]` does not alias the field), `field[i][k] = x` writes through;
  * `ti.template()` arguments are by reference: scalars live in a mutable `Box`, vectors / matrices /
    structs are mutable objects and augmented assignment updates them in place;
  * parallel for-loops run serially in index order, atomics return the old value.

What this cannot reproduce: Taichi's code generation (fast-math reassociation / fma contraction, its
pow / inverse lowering) and the nondeterministic order of its parallel atomics.  Fixtures made with it pin
the oracle's restatement of the ALGORITHM to float tolerance, not bit for bit.
"""
import ast
import functools
import inspect
import itertools
import textwrap

import os

import numpy as np

# TI_SHIM_FMA=1: accumulate dot products / squared norms with fused multiply-adds, as a fast-math code generator may
# (sensitivity study of the emulation caveat; the committed fixtures are made WITHOUT it)
_FMA = os.environ.get("TI_SHIM_FMA") == "1"


def _mac(s, a, b):
    """s + a * b in f32: separately rounded, or fused (exact product in f64, one rounding) under TI_SHIM_FMA."""
    if _FMA:
        return np.float32(np.float64(s) + np.float64(a) * np.float64(b))
    return s + a * b


f32 = np.float32
f64 = np.float64
i32 = np.int32
i64 = np.int64
u8 = np.uint8
gpu = cpu = cuda = vulkan = "cpu"


def init(*args, **kwargs):
    return None


def data_oriented(cls):
    return cls


def template():
    return "template"


def static(x):
    return x


class _Types:
    @staticmethod
    def ndarray(*a, **k):
        return "ndarray"

    @staticmethod
    def vector(n, dtype=float):
        return ("vector", n, dtype)

    @staticmethod
    def matrix(n, m, dtype=float):
        return ("matrix", n, m, dtype)


types = _Types()


def _np_dtype(dt):
    if dt in (float, f32, "f32"):
        return np.float32
    if dt in (int, i32, "i32"):
        return np.int32
    if dt is f64:
        return np.float64
    if dt is i64:
        return np.int64
    return np.dtype(dt).type


# --------------------------------------------------------------------------------------------- values
class Box:
    """A kernel-scope scalar variable (mutable so that ti.template() arguments are by reference)."""
    __slots__ = ("v",)
    __array_ufunc__ = None

    def __init__(self, v):
        self.v = v

    def _set(self, r):
        r = _u(r)
        if isinstance(self.v, (int, np.integer)) and not isinstance(self.v, (bool, np.bool_)):
            self.v = int(r)          # Taichi: the variable keeps its declared type (float -> int truncates)
        elif isinstance(self.v, np.floating):
            self.v = np.float32(r)
        else:
            self.v = r
        return self

    def __index__(self):
        return int(self.v)

    __int__ = __index__

    def __float__(self):
        return float(self.v)

    def __bool__(self):
        return bool(self.v)

    def __hash__(self):
        return hash(self.v)

    def __repr__(self):
        return f"Box({self.v!r})"

    def __neg__(self):
        return _c(-self.v)

    def __pos__(self):
        return _c(self.v)

    def __abs__(self):
        return _c(abs(self.v))


def _binop(name):
    import operator
    op = getattr(operator, name)

    def fwd(self, other):
        return _c(op(self.v, _u(other)))

    def rev(self, other):
        return _c(op(_u(other), self.v))

    def inplace(self, other):
        return self._set(op(self.v, _u(other)))

    return fwd, rev, inplace


for _n, _sym in (("add", "add"), ("sub", "sub"), ("mul", "mul"), ("truediv", "truediv"), ("floordiv", "floordiv"),
                 ("mod", "mod"), ("pow", "pow")):
    _f, _r, _i = _binop(_n)
    setattr(Box, f"__{_sym}__", _f)
    setattr(Box, f"__r{_sym}__", _r)
    setattr(Box, f"__i{_sym}__", _i)
for _n in ("lt", "le", "gt", "ge", "eq", "ne"):
    def _cmp(self, other, _op=getattr(__import__("operator"), _n)):
        return bool(_op(self.v, _u(other)))
    setattr(Box, f"__{_n}__", _cmp)


def _u(x):
    """Unbox."""
    return x.v if isinstance(x, Box) else x


def _c(x):
    """Canonical kernel-scope value: floats are f32, numpy ints are Python ints, the rest passes."""
    if isinstance(x, (float, np.floating)):
        return np.float32(x)
    if isinstance(x, np.integer):
        return int(x)
    return x


def _assign(x):
    """`name = value` in kernel scope: a fresh variable holding a copy."""
    if isinstance(x, Box):
        return Box(x.v)
    if isinstance(x, (Vector, Matrix)):
        return x.copy()
    if isinstance(x, (bool, np.bool_)):
        return bool(x)
    if isinstance(x, (float, np.floating)):
        return Box(np.float32(x))
    if isinstance(x, (int, np.integer)):
        return Box(int(x))
    return x


def _arr(x):
    if isinstance(x, (Vector, Matrix)):
        return x.a
    x = _u(x)
    if isinstance(x, (float, np.floating)):
        return np.float32(x)
    return x


class Vector:
    """ti.Vector value; `a` may be a view into a field (writes go through)."""
    __slots__ = ("a",)
    __array_ufunc__ = None

    def __init__(self, data, dt=None):
        if isinstance(data, np.ndarray) and dt is None and data.dtype in (np.float32, np.int32, np.int64):
            self.a = data
            return
        vals = [_u(v) for v in (data.a if isinstance(data, Vector) else data)]
        if dt is not None:
            self.a = np.array(vals, dtype=_np_dtype(dt))
        elif all(isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)) for v in vals):
            self.a = np.array(vals, dtype=np.int64)
        else:
            self.a = np.array(vals, dtype=np.float32)

    # constructors
    @staticmethod
    def zero(dt, n):
        return Vector(np.zeros(n, dtype=_np_dtype(dt)))

    @staticmethod
    def field(n, dtype=float, shape=()):
        return Field(dtype, shape, (n,))

    def copy(self):
        return Vector(self.a.copy())

    def to_numpy(self):
        return self.a.copy()

    def __len__(self):
        return self.a.shape[0]

    def __iter__(self):
        return (_c(v) for v in self.a)

    def __getitem__(self, i):
        return _c(self.a[int(_u(i))])

    def __setitem__(self, i, v):
        self.a[int(_u(i))] = _arr(v)

    def __repr__(self):
        return f"Vector({self.a.tolist()})"

    def _res(self, r):
        if r.dtype == np.float64:   # int vector with a weak Python float etc.
            r = r.astype(np.float32)
        return Vector(r)

    def __neg__(self):
        return Vector(-self.a)

    def cast(self, dt):
        if _np_dtype(dt) in (np.int32, np.int64):
            return Vector(np.trunc(self.a).astype(np.int64))
        return Vector(self.a.astype(np.float32))

    def norm_sqr(self):
        s = self.a[0] * self.a[0]
        for k in range(1, self.a.shape[0]):
            s = _mac(s, self.a[k], self.a[k])
        return _c(s)

    def norm(self):
        return np.sqrt(np.float32(self.norm_sqr()))

    def dot(self, o):
        b = _arr(o)
        s = self.a[0] * b[0]
        for k in range(1, self.a.shape[0]):
            s = _mac(s, self.a[k], b[k])
        return _c(s)

    def cross(self, o):
        a, b = self.a, _arr(o)
        return Vector(np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]],
                               dtype=np.float32))

    def outer_product(self, o):
        b = _arr(o)
        return Matrix(np.array([[x * y for y in b] for x in self.a], dtype=np.float32))


def _vec_ops(cls):
    import operator
    for name in ("add", "sub", "mul", "truediv"):
        op = getattr(operator, name)

        def fwd(self, other, op=op):
            return self._res(op(self.a, _arr(other)))

        def rev(self, other, op=op):
            return self._res(op(_arr(other), self.a))

        def inplace(self, other, op=op):
            r = op(self.a, _arr(other))
            self.a[...] = np.trunc(r) if self.a.dtype.kind == "i" and r.dtype.kind == "f" else r
            return self

        setattr(cls, f"__{name}__", fwd)
        setattr(cls, f"__r{name}__", rev)
        setattr(cls, f"__i{name}__", inplace)


_vec_ops(Vector)


class Matrix:
    __slots__ = ("a",)
    __array_ufunc__ = None

    def __init__(self, data, dt=None):
        if isinstance(data, np.ndarray) and data.dtype == np.float32 and dt is None:
            self.a = data
        else:
            self.a = np.array([[_u(v) for v in row] for row in (data.a if isinstance(data, Matrix) else data)],
                              dtype=np.float32)

    @staticmethod
    def zero(dt, n, m):
        return Matrix(np.zeros((n, m), dtype=np.float32))

    @staticmethod
    def identity(dt, n):
        return Matrix(np.eye(n, dtype=np.float32))

    @staticmethod
    def field(n, m, dtype=float, shape=()):
        return Field(dtype, shape, (n, m))

    def copy(self):
        return Matrix(self.a.copy())

    def to_numpy(self):
        return self.a.copy()

    def _res(self, r):
        return Matrix(r.astype(np.float32) if r.dtype != np.float32 else r)

    def __neg__(self):
        return Matrix(-self.a)

    def __getitem__(self, ij):
        i, j = ij
        return _c(self.a[int(_u(i)), int(_u(j))])

    def __setitem__(self, ij, v):
        i, j = ij
        self.a[int(_u(i)), int(_u(j))] = _arr(v)

    def __matmul__(self, o):
        n, m = self.a.shape
        if isinstance(o, Vector):
            out = np.zeros(n, dtype=np.float32)
            for i in range(n):
                s = self.a[i, 0] * o.a[0]
                for k in range(1, m):
                    s = s + self.a[i, k] * o.a[k]
                out[i] = s
            return Vector(out)
        b = o.a
        out = np.zeros((n, b.shape[1]), dtype=np.float32)
        for i in range(n):
            for j in range(b.shape[1]):
                s = self.a[i, 0] * b[0, j]
                for k in range(1, m):
                    s = s + self.a[i, k] * b[k, j]
                out[i, j] = s
        return Matrix(out)

    def transpose(self):
        return Matrix(self.a.T.copy())

    def __repr__(self):
        return f"Matrix({self.a.tolist()})"


_vec_ops(Matrix)


class Struct:
    def __init__(self, **kw):
        for k, v in kw.items():
            v = _u(v)
            setattr(self, k, v.copy() if isinstance(v, (Vector, Matrix)) else _c(v))


# --------------------------------------------------------------------------------------------- fields
class Field:
    """ti.field / ti.Vector.field / ti.Matrix.field over one numpy array."""

    def __init__(self, dtype, shape, elem=()):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.elem = tuple(elem)
        self.arr = np.zeros(self.shape + self.elem, dtype=_np_dtype(dtype))

    def _idx(self, idx):
        if idx is None:
            return ()
        if isinstance(idx, Vector):
            idx = tuple(int(v) for v in idx.a)
        elif isinstance(idx, tuple):
            idx = tuple(int(_u(v)) for v in idx)
        else:
            idx = (int(_u(idx)),)
        for k, n in zip(idx, self.shape):
            if not 0 <= k < n:
                raise IndexError(f"field index {idx} out of range for shape {self.shape} (Taichi would read out of bounds)")
        return idx

    def __getitem__(self, idx):
        idx = self._idx(idx)
        if len(self.elem) == 0:
            return _c(self.arr[idx])
        if len(self.elem) == 1:
            return Vector(self.arr[idx])
        return Matrix(self.arr[idx])

    def __setitem__(self, idx, v):
        idx = self._idx(idx)
        v = _arr(v)
        if self.arr.dtype.kind == "i" and isinstance(v, (float, np.floating)):
            v = int(v)
        self.arr[idx] = v

    def fill(self, v):
        self.arr[...] = _arr(v)

    def to_numpy(self):
        return self.arr.copy()

    def from_numpy(self, a):
        self.arr[...] = a


def field(dtype=float, shape=()):
    return Field(dtype, shape)


class _Math:
    pi = np.pi

    @staticmethod
    def dot(a, b):
        return a.dot(b)

    @staticmethod
    def cross(a, b):
        return a.cross(b)

    @staticmethod
    def inverse(m):
        a = m.a
        if a.shape == (2, 2):
            det = a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
            inv = np.float32(1.0) / det
            return Matrix(np.array([[a[1, 1] * inv, -a[0, 1] * inv], [-a[1, 0] * inv, a[0, 0] * inv]], dtype=np.float32))
        # 3x3: adjugate / determinant, all in f32
        c = lambda i, j: (a[(i + 1) % 3, (j + 1) % 3] * a[(i + 2) % 3, (j + 2) % 3]
                          - a[(i + 1) % 3, (j + 2) % 3] * a[(i + 2) % 3, (j + 1) % 3])
        det = a[0, 0] * c(0, 0) + a[0, 1] * c(0, 1) + a[0, 2] * c(0, 2)
        inv = np.float32(1.0) / det
        out = np.zeros((3, 3), dtype=np.float32)
        for i in range(3):
            for j in range(3):
                out[j, i] = c(i, j) * inv
        return Matrix(out)


math = _Math()


def max(a, b):   # noqa: A001  (Taichi's name)
    a, b = _c(_u(a)), _c(_u(b))
    return a if a >= b else b


def min(a, b):   # noqa: A001
    a, b = _c(_u(a)), _c(_u(b))
    return a if a <= b else b


def abs(a):      # noqa: A001
    a = _u(a)
    return -a if a < 0 else a


def sqrt(a):
    return np.sqrt(np.float32(_u(a)))


def pow(a, b):   # noqa: A001
    return np.power(np.float32(_u(a)), np.float32(_u(b)))


def cast(v, dt):
    v = _u(v)
    if isinstance(v, (Vector, Matrix)):
        return v.cast(dt)
    return int(v) if _np_dtype(dt) in (np.int32, np.int64) else np.float32(v)


def floor(a):
    return np.floor(np.float32(_u(a)))


def ndrange(*ranges):
    its = [range(*r) if isinstance(r, tuple) else range(int(_u(r))) for r in ranges]
    return itertools.product(*its)


def grouped(x):
    if isinstance(x, Field):
        return (Vector(np.array(i, dtype=np.int64)) for i in np.ndindex(*x.shape))
    return (Vector(np.array(i, dtype=np.int64)) for i in x)


def _atomic(container, index, val, sign):
    old = container[index]
    container[index] = old + sign * _u(val)
    return old


class _PrefixSum:
    def __init__(self, n):
        self.n = n

    def run(self, f):
        f.arr[...] = np.cumsum(f.arr, dtype=np.int64).astype(f.arr.dtype)   # inclusive, like ti.algorithms


class _Algorithms:
    PrefixSumExecutor = _PrefixSum


algorithms = _Algorithms()


# --------------------------------------------------------------------------------------------- kernels
class _Rewrite(ast.NodeTransformer):
    """Kernel-scope semantics by source rewriting (see the module docstring)."""

    def visit_FunctionDef(self, node):
        node.decorator_list = []
        node.returns = None
        for a in node.args.args + node.args.kwonlyargs:
            a.annotation = None
        self.generic_visit(node)
        return node

    @staticmethod
    def _call(name, *args):
        return ast.Call(func=ast.Name(id=name, ctx=ast.Load()), args=list(args), keywords=[])

    def _inner(self, node):
        """Visit an attribute chain / field expression that is only a stepping stone (no canonicalisation)."""
        if isinstance(node, ast.Attribute):
            node.value = self._inner(node.value)
            return node
        return self.visit(node)

    def visit_Attribute(self, node):
        node.value = self._inner(node.value)
        return self._call("_ti_c", node) if isinstance(node.ctx, ast.Load) else node

    def visit_Subscript(self, node):
        node.value = self._inner(node.value) if isinstance(node.value, ast.Attribute) else self.visit(node.value)
        node.slice = self.visit(node.slice)
        return self._call("_ti_c", node) if isinstance(node.ctx, ast.Load) else node

    def visit_Constant(self, node):
        return self._call("_ti_c", node) if isinstance(node.value, float) else node

    def visit_Call(self, node):
        f = node.func
        if (isinstance(f, ast.Attribute) and f.attr in ("atomic_add", "atomic_sub") and isinstance(f.value, ast.Name)
                and f.value.id == "ti" and isinstance(node.args[0], ast.Subscript)):
            tgt = node.args[0]
            sign = ast.Constant(value=1 if f.attr == "atomic_add" else -1)
            return self._call("_ti_atomic", self._inner(tgt.value), self.visit(tgt.slice), self.visit(node.args[1]), sign)
        node.func = self._inner(f) if isinstance(f, ast.Attribute) else self.visit(f)
        node.args = [self.visit(a) for a in node.args]
        node.keywords = [ast.keyword(arg=k.arg, value=self.visit(k.value)) for k in node.keywords]
        return node

    def visit_Assign(self, node):
        self.generic_visit(node)
        if all(isinstance(t, ast.Name) for t in node.targets):
            node.value = self._call("_ti_assign", node.value)
        return node


_HELPERS = {"_ti_c": _c, "_ti_assign": _assign, "_ti_atomic": _atomic}


def _compile(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    tree = _Rewrite().visit(tree)
    ast.fix_missing_locations(tree)
    ns = dict(fn.__globals__)
    ns.update(_HELPERS)
    exec(compile(tree, inspect.getsourcefile(fn) or "<taichi-shim>", "exec"), ns)
    return ns[fn.__name__]


def _lazy(fn, is_kernel):
    state = {}

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        impl = state.get("impl")
        if impl is None:
            impl = state["impl"] = _compile(fn)
        if not is_kernel:
            return impl(*args, **kwargs)
        args = [np.float32(a) if isinstance(a, float) else a for a in args]
        out = impl(*args, **kwargs)
        out = _u(out)
        if isinstance(out, np.floating):
            return float(out)
        return out

    wrapper.__ti_kind__ = "kernel" if is_kernel else "func"
    return wrapper


def kernel(fn):
    return _lazy(fn, True)


def func(fn):
    return _lazy(fn, False)


def __getattr__(name):   # anything else the unused solvers (PBF, IISPH, shape matching) mention at import time
    def _missing(*a, **k):
        raise NotImplementedError(f"taichi shim: ti.{name} is not emulated")
    return _missing
