"""Minimal `jax` stand-in so that the reference's UNMODIFIED source can be executed here.

TEST INFRASTRUCTURE ONLY (see oracle/ref_shim/README.md).  JAX / jaxlib are not installable
in this image (no wheel, no network).  The reference's hot path (mlift/systems.py,
solvers.py, linalg.py, ode.py, prior.py, distributions.py, transforms.py, math.py and the
example models) is plain `jax.numpy` + `lax.scan / while_loop / cond` + the function
transforms `jit / jvp / jacfwd / jacrev / grad / value_and_grad / vmap`.  This package
maps those names onto eager torch-fp64 arithmetic:

  * arrays       -> `Array`, an immutable wrapper over a torch tensor (`_array.py`)
  * jit          -> identity (NumPy arguments become `Array`s, as `jit` would)
  * scan / while_loop / cond -> Python loops / branches (`lax/`)
  * jvp / grad / value_and_grad -> `torch.func.jvp / grad` (composable, so the reference's
    reverse-over-forward gradient of the log-determinant runs as written)
  * jacfwd       -> one `torch.func.jvp` per input direction (no `vmap`, so value-dependent
    `lax.cond`s inside the differentiated function keep working)
  * vmap         -> Python loop + stack

What this changes relative to the real thing: XLA's fused kernels / LAPACK calls become
torch's (LAPACK via torch.linalg), i.e. rounding-level differences only; the algorithm is
the reference's own source, line for line.
"""
import numpy as onp
import torch

from . import config as _config_module
from ._array import Array, as_f, as_t, tree_leaves, tree_map, unwrap, wrap
from . import numpy  # noqa: F401  (jax.numpy)
from . import lax  # noqa: F401
from . import scipy  # noqa: F401
from . import core  # noqa: F401

config = _config_module
__version__ = "0.4.14+torch-shim"


def jit(fun=None, static_argnames=None, static_argnums=None, **_ignored):
    if fun is None:
        return lambda f: jit(f, static_argnames, static_argnums)

    def jitted(*args, **kwargs):
        return wrap(fun(*wrap(args), **wrap(kwargs)))  # outputs are always jax arrays

    jitted.__wrapped__ = fun
    jitted.__name__ = getattr(fun, "__name__", "jitted")
    return jitted


def _call(fun, args, kwargs, argnum, x):
    a = list(args)
    a[argnum] = wrap(x)
    return fun(*a, **kwargs)


def jvp(fun, primals, tangents, has_aux=False):
    p = tuple(tree_map(as_f, x) for x in primals)
    t = tuple(tree_map(as_f, x) for x in tangents)

    def g(*xs):
        return unwrap(fun(*wrap(xs)))

    out, tout = torch.func.jvp(g, p, t)
    return wrap(out), wrap(tout)


def jacfwd(fun, argnums=0, has_aux=False):
    def jac(*args, **kwargs):
        x = as_f(args[argnums])

        def g(xx):
            return unwrap(_call(fun, args, kwargs, argnums, xx))

        try:  # all directions in one pass (vmap over jvp)
            out = torch.func.jacfwd(g)(x)
        except Exception:  # an op without a batching rule: one jvp per direction
            cols = []
            for k in range(x.numel()):
                e = torch.zeros(x.numel(), dtype=x.dtype)
                e[k] = 1.0
                cols.append(torch.func.jvp(g, (x,), (e.reshape(x.shape),))[1])
            out = tree_map(
                lambda *c: torch.stack(c, -1).reshape(tuple(c[0].shape) + tuple(x.shape)),
                cols[0], *cols[1:])  # fmt: skip
        return wrap(out)

    return jac


def jacrev(fun, argnums=0, has_aux=False):
    def jac(*args, **kwargs):
        x = as_f(args[argnums])

        def g(xx):
            return as_t(_call(fun, args, kwargs, argnums, xx))

        out, vjp_fn = torch.func.vjp(g, x)
        rows = []
        for k in range(out.numel()):
            e = torch.zeros(out.numel(), dtype=out.dtype)
            e[k] = 1.0
            rows.append(vjp_fn(e.reshape(out.shape))[0])
        return Array(torch.stack(rows).reshape(tuple(out.shape) + tuple(x.shape)))

    return jac


def value_and_grad(fun, argnums=0, has_aux=False):
    def vg(*args, **kwargs):
        x = as_f(args[argnums])

        def g(xx):
            out = _call(fun, args, kwargs, argnums, xx)
            if has_aux:
                val, aux = out
                return as_t(val), (as_t(val), unwrap(aux))
            return as_t(out), (as_t(out), ())

        gr, (val, aux) = torch.func.grad(g, has_aux=True)(x)
        if has_aux:
            return (Array(val), wrap(aux)), Array(gr)
        return Array(val), Array(gr)

    return vg


def grad(fun, argnums=0, has_aux=False):
    vg = value_and_grad(fun, argnums, has_aux)

    def gf(*args, **kwargs):
        v, g = vg(*args, **kwargs)
        return (g, v[1]) if has_aux else g

    return gf


def vmap(fun, in_axes=0, out_axes=0):
    def vf(*args):
        axes = list(in_axes) if isinstance(in_axes, (tuple, list)) else [in_axes] * len(args)
        axes += [None] * (len(args) - len(axes))
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = tree_leaves(wrap(a))[0].shape[ax]
                break

        def take(a, ax, i):
            if ax is None:
                return a
            return tree_map(lambda l: Array(as_t(l).select(ax, i)), wrap(a))

        outs = [fun(*[take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]

        def stack(ax, first, *rest):
            if ax is None:  # unbatched output: every instance returns the same value
                return first
            return tree_map(
                lambda f, *r: Array(torch.stack([as_t(f)] + [as_t(x) for x in r], ax)),
                first, *rest)  # fmt: skip

        if isinstance(out_axes, (tuple, list)):  # one axis (or None) per element of the output
            return tuple(stack(ax, outs[0][k], *[o[k] for o in outs[1:]])
                         for k, ax in enumerate(out_axes))
        return stack(out_axes, outs[0], *outs[1:])

    return vf


def device_get(x):
    return tree_map(lambda l: onp.asarray(l) if isinstance(l, Array) else l, x)
