"""Immutable array type of the `jax` stand-in: a thin wrapper over a torch fp64 tensor.

TEST INFRASTRUCTURE ONLY (see oracle/ref_shim/README.md).  JAX arrays are immutable, so
`x += y` inside the reference rebinds `x`; a bare torch tensor would be mutated in place
and alias the caller's array.  `Array` defines no in-place operators, which gives the
reference's source exactly the semantics it was written for, and keeps whatever tensor
subclass `torch.func` hands in (dual / grad-tracking wrappers), so the transforms compose.
"""
import builtins

import numpy as onp
import torch

F64 = torch.float64


def as_t(x):
    """Anything array-like -> torch tensor (fp64 for floats, int64 for ints)."""
    if isinstance(x, Array):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, onp.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, onp.integer)):
        return torch.tensor(int(x), dtype=torch.int64)
    if isinstance(x, (float, onp.floating)):
        return torch.tensor(float(x), dtype=F64)
    if isinstance(x, onp.ndarray):
        if x.dtype == onp.float32:
            x = x.astype(onp.float64)
        return torch.from_numpy(onp.array(x, order="C", copy=True))  # keeps 0-d as 0-d
    if isinstance(x, (list, tuple)):
        if len(x) == 0:
            return torch.zeros(0, dtype=F64)
        parts = [as_t(e) for e in x]
        if any(p.is_floating_point() for p in parts):
            parts = [p.to(F64) for p in parts]
        return torch.stack(parts)
    raise TypeError(f"cannot convert {type(x)} to an array")


def as_f(x):
    """as_t, then promoted to fp64 (elementwise transcendental functions)."""
    t = as_t(x)
    return t if t.is_floating_point() else t.to(F64)


def _idx(i):
    if isinstance(i, Array):
        return i.t
    if isinstance(i, onp.ndarray):
        return torch.from_numpy(i.copy())
    if isinstance(i, tuple):
        return tuple(_idx(j) for j in i)
    return i


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        t = self.arr.t.clone()
        t[_idx(self.idx)] = as_t(v).to(t.dtype)
        return Array(t)

    def add(self, v):
        t = self.arr.t.clone()
        t[_idx(self.idx)] = t[_idx(self.idx)] + as_t(v).to(t.dtype)
        return Array(t)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


def _binop(fn, reflected=False):
    def op(self, other):
        if isinstance(other, (dict, str)) or other is None:
            return NotImplemented
        a = self.t
        if type(other) in (int, float, bool):
            # keep Python scalars scalar: torch.pow(Tensor, Scalar) has no derivative in
            # the exponent (a tensor exponent gives 0 * log(negative base) = NaN tangents)
            if type(other) is float and not a.is_floating_point():
                a = a.to(F64)
            b = other
        else:
            b = as_t(other)
        if reflected:
            a, b = b, a
        return Array(fn(a, b))

    return op


def _is_f(x):
    return x.is_floating_point() if isinstance(x, torch.Tensor) else isinstance(x, float)


def _pow(a, b):
    if isinstance(a, torch.Tensor) and not a.is_floating_point() and _is_f(b):
        a = a.to(F64)
    return torch.pow(a, b)


def _truediv(a, b):
    if isinstance(a, torch.Tensor) and not a.is_floating_point():
        a = a.to(F64)
    if isinstance(b, torch.Tensor) and not b.is_floating_point():
        b = b.to(F64)
    return a / b


def _sub(a, b):
    return a - b


def _mod(a, b):
    return a % b


def _matmul(a, b):
    if a.is_floating_point() != b.is_floating_point():
        a, b = a.to(F64), b.to(F64)
    return a @ b


class Array:
    __slots__ = ("t",)
    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """`ndarray <op> Array` defers to the operators below (the result stays an
        `Array`, as with a jax array); any other ufunc is evaluated by NumPy on the
        converted values (e.g. `onp.isnan(error)` in mlift/solvers.py:86)."""
        if method == "__call__" and len(inputs) == 2 and not kwargs and ufunc in _UFUNC_OPS:
            a, b = inputs
            op, rop = _UFUNC_OPS[ufunc]
            return getattr(a, op)(b) if isinstance(a, Array) else getattr(b, rop)(a)
        args = [onp.asarray(x) if isinstance(x, Array) else x for x in inputs]
        return getattr(ufunc, method)(*args, **kwargs)

    def __init__(self, t):
        self.t = t

    # -- introspection
    shape = property(lambda self: tuple(self.t.shape))
    ndim = property(lambda self: self.t.dim())
    size = property(lambda self: self.t.numel())
    T = property(lambda self: Array(self.t.T if self.t.dim() == 2 else self.t))
    at = property(lambda self: _At(self))

    @property
    def dtype(self):
        return onp.dtype(
            {torch.float64: "float64", torch.int64: "int64", torch.bool: "bool",
             torch.float32: "float32", torch.int32: "int32"}[self.t.dtype])  # fmt: skip

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return (Array(self.t[i]) for i in range(self.t.shape[0]))

    def __getitem__(self, i):
        return Array(self.t[_idx(i)])

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.t)

    def __int__(self):
        return int(self.t)

    def __index__(self):
        return int(self.t)

    def __bool__(self):
        return bool(self.t)

    def __format__(self, spec):
        return format(self.t.item(), spec) if self.t.dim() == 0 else repr(self)

    def __repr__(self):
        return f"Array({onp.asarray(self)!r})"

    def item(self):
        return self.t.item()

    # -- arithmetic (no in-place forms on purpose)
    __add__ = _binop(lambda a, b: a + b)
    __radd__ = _binop(lambda a, b: a + b, True)
    __sub__ = _binop(_sub)
    __rsub__ = _binop(_sub, True)
    __mul__ = _binop(lambda a, b: a * b)
    __rmul__ = _binop(lambda a, b: a * b, True)
    __truediv__ = _binop(_truediv)
    __rtruediv__ = _binop(_truediv, True)
    __pow__ = _binop(_pow)
    __rpow__ = _binop(_pow, True)
    __matmul__ = _binop(_matmul)
    __rmatmul__ = _binop(_matmul, True)
    __floordiv__ = _binop(lambda a, b: a // b)
    __mod__ = _binop(_mod)
    __lt__ = _binop(lambda a, b: a < b)
    __le__ = _binop(lambda a, b: a <= b)
    __gt__ = _binop(lambda a, b: a > b)
    __ge__ = _binop(lambda a, b: a >= b)
    __eq__ = _binop(lambda a, b: a == b)
    __ne__ = _binop(lambda a, b: a != b)
    __and__ = _binop(lambda a, b: torch.logical_and(a, torch.as_tensor(b)))
    __or__ = _binop(lambda a, b: torch.logical_or(a, torch.as_tensor(b)))
    __hash__ = None

    def __neg__(self):
        return Array(-self.t)

    def __pos__(self):
        return self

    def __abs__(self):
        return Array(self.t.abs())

    def __invert__(self):
        return Array(torch.logical_not(self.t))

    # -- methods the reference calls on arrays
    def sum(self, axis=None):
        return Array(self.t.sum() if axis is None else self.t.sum(axis))

    def prod(self, axis=None):
        return Array(self.t.prod() if axis is None else self.t.prod(axis))

    def mean(self, axis=None):
        return Array(self.t.mean() if axis is None else self.t.mean(axis))

    def max(self, axis=None):
        if axis is not None:
            return Array(self.t.max(axis).values)
        # NaN-propagating like numpy / XLA reduce-max
        return Array(self.t.max())

    def min(self, axis=None):
        return Array(self.t.min() if axis is None else self.t.min(axis).values)

    def cumsum(self, axis=0):
        return Array(self.t.cumsum(axis))

    def diagonal(self):
        return Array(self.t.diagonal())

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Array(self.t.reshape(shape))

    def flatten(self):
        return Array(self.t.reshape(-1))

    def astype(self, dtype):
        return Array(self.t.to(_torch_dtype(dtype)))

    def copy(self):
        return Array(self.t.clone())

    def all(self):
        return Array(self.t.all())

    def any(self):
        return Array(self.t.any())

    def squeeze(self):
        return Array(self.t.squeeze())


_UFUNC_OPS = {
    onp.add: ("__add__", "__radd__"), onp.subtract: ("__sub__", "__rsub__"),
    onp.multiply: ("__mul__", "__rmul__"), onp.true_divide: ("__truediv__", "__rtruediv__"),
    onp.power: ("__pow__", "__rpow__"), onp.matmul: ("__matmul__", "__rmatmul__"),
    onp.less: ("__lt__", "__gt__"), onp.less_equal: ("__le__", "__ge__"),
    onp.greater: ("__gt__", "__lt__"), onp.greater_equal: ("__ge__", "__le__"),
}  # fmt: skip


def _torch_dtype(dtype):
    if dtype is None:
        return F64
    if dtype in (float, onp.float64, "float64") or dtype == onp.dtype("float64"):
        return F64
    if dtype in (int, onp.int64, "int64") or dtype == onp.dtype("int64"):
        return torch.int64
    if dtype in (builtins.bool, onp.bool_, "bool"):
        return torch.bool
    raise TypeError(f"unsupported dtype {dtype}")


# ---- pytrees (tuple / list / dict / None / leaves), as far as the reference uses them
def tree_map(fn, tree, *rest, is_leaf=None):
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*(tree_map(fn, a, *(r[i] for r in rest))
                            for i, a in enumerate(tree)))  # fmt: skip
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(fn, a, *(r[i] for r in rest)) for i, a in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(fn, v, *(r[k] for r in rest)) for k, v in tree.items()}
    if tree is None:
        return None
    return fn(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(out.append, tree)
    return out


def wrap(tree):
    """torch tensors / numpy arrays -> Array at the leaves; other leaves unchanged."""

    def f(x):
        if isinstance(x, torch.Tensor):
            return Array(x)
        if isinstance(x, onp.ndarray) and x.dtype.kind in "fiub":
            return Array(as_t(x))
        return x

    return tree_map(f, tree)


def unwrap(tree):
    """Array leaves -> torch tensors (python / numpy scalars become 0-d tensors)."""

    def f(x):
        if isinstance(x, (Array, onp.ndarray, float, int, onp.floating, onp.integer)):
            return as_t(x)
        return x

    return tree_map(f, tree)
