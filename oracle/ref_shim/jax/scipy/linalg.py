"""`jax.scipy.linalg` stand-in (torch fp64)."""
import torch

from .._array import Array, as_f


def cho_solve(c_and_lower, b):
    c, lower = c_and_lower
    c, b = as_f(c), as_f(b)
    vec = b.dim() == 1
    x = torch.cholesky_solve(b[:, None] if vec else b, c, upper=not lower)
    return Array(x[:, 0] if vec else x)


def cho_factor(a, lower=False):
    l = torch.linalg.cholesky(as_f(a))
    return (Array(l if lower else l.mT), lower)


def solve(a, b, assume_a="gen", sym_pos=False, lower=False):
    a, b = as_f(a), as_f(b)
    if sym_pos or assume_a == "pos":
        l = torch.linalg.cholesky(a)
        vec = b.dim() == 1
        x = torch.cholesky_solve(b[:, None] if vec else b, l)
        return Array(x[:, 0] if vec else x)
    return Array(torch.linalg.solve(a, b))


def solve_triangular(a, b, lower=False, trans=0):
    a, b = as_f(a), as_f(b)
    upper = not lower
    if trans in (1, "T"):
        a, upper = a.mT, not upper
    vec = b.dim() == 1
    x = torch.linalg.solve_triangular(a, b[:, None] if vec else b, upper=upper)
    return Array(x[:, 0] if vec else x)


def expm(a):
    return Array(torch.linalg.matrix_exp(as_f(a)))
