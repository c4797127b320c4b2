"""`jax.numpy` stand-in (torch fp64 backed): only what the reference's hot path calls.

TEST INFRASTRUCTURE ONLY (see oracle/ref_shim/README.md).
"""
import builtins

import numpy as onp
import torch

from .._array import F64, Array, _torch_dtype, as_f, as_t
from . import linalg  # noqa: F401

pi = onp.pi
inf = onp.inf
nan = onp.nan
float64 = onp.float64
int64 = onp.int64
newaxis = None
ndarray = Array
load = onp.load
save = onp.save
loadtxt = onp.loadtxt
finfo = onp.finfo
seterr = onp.seterr


def _unary(fn):
    def f(x):
        return Array(fn(as_f(x)))

    return f


log = _unary(torch.log)
exp = _unary(torch.exp)
expm1 = _unary(torch.expm1)
log1p = _unary(torch.log1p)
sqrt = _unary(torch.sqrt)
sin = _unary(torch.sin)
cos = _unary(torch.cos)
tan = _unary(torch.tan)
arctan = _unary(torch.arctan)
arccos = _unary(torch.arccos)
arcsin = _unary(torch.arcsin)
sinh = _unary(torch.sinh)
cosh = _unary(torch.cosh)
tanh = _unary(torch.tanh)
sign = _unary(torch.sign)
abs = absolute = _unary(torch.abs)
square = _unary(torch.square)
isnan = _unary(torch.isnan)
isfinite = _unary(torch.isfinite)


def logical_not(x):
    return Array(torch.logical_not(as_t(x)))


def _binary(fn):
    def f(a, b):
        return Array(fn(as_t(a), as_t(b)))

    return f


logical_or = _binary(torch.logical_or)
logical_and = _binary(torch.logical_and)
less = _binary(torch.lt)
less_equal = _binary(torch.le)
greater = _binary(torch.gt)
greater_equal = _binary(torch.ge)
maximum = _binary(torch.maximum)
minimum = _binary(torch.minimum)


def isscalar(x):
    return onp.isscalar(x)


def array(x, dtype=None):
    t = as_t(x)
    if dtype is not None:
        t = t.to(_torch_dtype(dtype))
    return Array(t)


def asarray(x, dtype=None):
    return array(x, dtype)


def zeros(shape, dtype=None):
    return Array(torch.zeros(shape, dtype=_torch_dtype(dtype)))


def ones(shape, dtype=None):
    return Array(torch.ones(shape, dtype=_torch_dtype(dtype)))


def zeros_like(x):
    return Array(torch.zeros_like(as_t(x)))


def ones_like(x):
    return Array(torch.ones_like(as_t(x)))


def full(shape, v):
    return Array(torch.full(shape if isinstance(shape, tuple) else (shape,), float(v),
                            dtype=F64))  # fmt: skip


def eye(n):
    return Array(torch.eye(n, dtype=F64))


identity = eye


def arange(*a):
    return Array(torch.arange(*a))


def linspace(start, stop, num=50):
    # host NumPy's linspace: it is what jax.numpy.linspace reproduces value for value
    return Array(torch.from_numpy(onp.linspace(float(start), float(stop), int(num))))


def diag(x, k=0):
    return Array(torch.diag(as_t(x), k))


def tril(x, k=0):
    return Array(torch.tril(as_t(x), k))


def triu(x, k=0):
    return Array(torch.triu(as_t(x), k))


def concatenate(parts, axis=0):
    ts = [as_t(p) for p in parts]
    if builtins.any(t.is_floating_point() for t in ts):
        ts = [t.to(F64) for t in ts]
    return Array(torch.cat(ts, axis))


def stack(parts, axis=0):
    ts = [as_t(p) for p in parts]
    if builtins.any(t.is_floating_point() for t in ts):
        ts = [t.to(F64) for t in ts]
    return Array(torch.stack(ts, axis))


def split(x, indices_or_sections, axis=0):
    t = as_t(x)
    if isinstance(indices_or_sections, int):
        return [Array(p) for p in torch.tensor_split(t, indices_or_sections, axis)]
    return [Array(p) for p in torch.tensor_split(t, list(indices_or_sections), axis)]


def where(c, a, b):
    a, b = as_t(a), as_t(b)
    if a.is_floating_point() != b.is_floating_point():
        a, b = a.to(F64), b.to(F64)
    return Array(torch.where(as_t(c), a, b))


def sum(x, axis=None):
    t = as_t(x)
    return Array(t.sum() if axis is None else t.sum(axis))


def prod(x, axis=None):
    t = as_t(x)
    return Array(t.prod() if axis is None else t.prod(axis))


def mean(x, axis=None):
    t = as_f(x)
    return Array(t.mean() if axis is None else t.mean(axis))


def max(x, axis=None):
    return Array(as_t(x)).max(axis)


def min(x, axis=None):
    return Array(as_t(x)).min(axis)


def cumsum(x, axis=0):
    return Array(as_t(x).cumsum(axis))


def diff(x):
    t = as_t(x)
    return Array(t[1:] - t[:-1])


def clip(x, lo, hi):
    return Array(torch.clamp(as_t(x), lo, hi))


def dot(a, b):
    return Array(as_t(a) @ as_t(b))


def outer(a, b):
    return Array(torch.outer(as_t(a), as_t(b)))


def reshape(x, shape):
    return Array(as_t(x).reshape(shape))


def atleast_1d(x):
    t = as_t(x)
    return Array(t.reshape(1) if t.dim() == 0 else t)


def atleast_2d(x):
    t = as_t(x)
    while t.dim() < 2:
        t = t.unsqueeze(0)
    return Array(t)


def pad(x, widths):
    t = as_t(x)
    if isinstance(widths, int):
        widths = (widths, widths)
    if t.dim() == 1:
        lo, hi = widths if not isinstance(widths[0], (tuple, list)) else widths[0]
        return Array(torch.nn.functional.pad(t, (lo, hi)))
    flat = []
    for lo, hi in reversed(list(widths)):
        flat += [lo, hi]
    return Array(torch.nn.functional.pad(t, tuple(flat)))
