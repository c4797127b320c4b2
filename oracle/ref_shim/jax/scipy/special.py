"""`jax.scipy.special` stand-in (torch fp64; all differentiable through torch.func)."""
import torch

from .._array import Array, as_f


def _unary(fn):
    def f(x):
        return Array(fn(as_f(x)))

    return f


ndtr = _unary(torch.special.ndtr)
ndtri = _unary(torch.special.ndtri)
log_ndtr = _unary(torch.special.log_ndtr)
gammaln = _unary(torch.special.gammaln)
expit = _unary(torch.special.expit)
logit = _unary(torch.special.logit)
erf = _unary(torch.special.erf)
erfc = _unary(torch.special.erfc)
