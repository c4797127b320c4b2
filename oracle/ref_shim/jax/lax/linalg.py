"""`jax.lax.linalg` stand-in: lower Cholesky factor and triangular solves, torch fp64."""
import torch

from .._array import Array, as_f


def cholesky(a, symmetrize_input=True):
    a = as_f(a)
    if symmetrize_input:
        a = (a + a.mT) / 2
    l, info = torch.linalg.cholesky_ex(a)
    if bool((info > 0).any()):  # XLA returns NaNs instead of raising
        l = torch.full_like(l, float("nan"))
    return Array(l)


def triangular_solve(a, b, left_side=False, lower=False, transpose_a=False,
                     conjugate_a=False, unit_diagonal=False):  # fmt: skip
    a, b = as_f(a), as_f(b)
    upper = not lower
    if transpose_a:
        a, upper = a.mT, not upper
    vec = b.dim() == 1
    if vec:
        b = b[:, None] if left_side else b[None, :]
    x = torch.linalg.solve_triangular(a, b, upper=upper, left=left_side,
                                      unitriangular=unit_diagonal)  # fmt: skip
    return Array(x[:, 0] if (vec and left_side) else x[0] if vec else x)
