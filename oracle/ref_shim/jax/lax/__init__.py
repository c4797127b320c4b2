"""`jax.lax` stand-in: structured control flow run as plain Python loops.

TEST INFRASTRUCTURE ONLY (see oracle/ref_shim/README.md).  `scan` / `while_loop` /
`cond` execute eagerly with the semantics of the JAX primitives (carry threading, stacked
outputs, `reverse=`); nothing is traced, so predicates are read as Python booleans.
"""
import torch

from .._array import Array, as_f, as_t, tree_leaves, tree_map
from . import linalg  # noqa: F401


def _unary(fn):
    def f(x):
        return Array(fn(as_f(x)))

    return f


log = _unary(torch.log)
exp = _unary(torch.exp)
expm1 = _unary(torch.expm1)
log1p = _unary(torch.log1p)
sqrt = _unary(torch.sqrt)


def gt(a, b):
    return Array(torch.gt(as_t(a), as_t(b)))


def lt(a, b):
    return Array(torch.lt(as_t(a), as_t(b)))


def cond(pred, true_fun, false_fun, *operands, operand="__unset__"):
    if operand != "__unset__":
        operands = (operand,)
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def _as_leaf(x):
    return x if isinstance(x, Array) else Array(as_t(x))


def scan(f, init, xs, length=None, reverse=False):
    if xs is None:
        n = int(length)
    else:
        xs = tree_map(_as_leaf, xs)
        n = tree_leaves(xs)[0].shape[0]
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        x = None if xs is None else tree_map(lambda a: a[i], xs)
        carry, ys[i] = f(carry, x)
    if n == 0 or ys[0] is None:
        return carry, None
    stacked = tree_map(
        lambda first, *rest: Array(torch.stack([as_t(first)] + [as_t(r) for r in rest])),
        ys[0], *ys[1:])  # fmt: skip
    return carry, stacked
