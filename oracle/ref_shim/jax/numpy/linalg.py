"""`jax.numpy.linalg` stand-in: general solve (pivoted LU) and slogdet, torch fp64."""
import torch

from .._array import Array, as_f


def solve(a, b):
    return Array(torch.linalg.solve(as_f(a), as_f(b)))


def slogdet(a):
    s, l = torch.linalg.slogdet(as_f(a))
    return Array(s), Array(l)


def cholesky(a):
    return Array(torch.linalg.cholesky(as_f(a)))


def inv(a):
    return Array(torch.linalg.inv(as_f(a)))
