"""Block algebra of `GaussianProcessModelSystem` (reference mlift/systems.py:859-918:
`y = cholesky(covar_func(u, data)) @ n`), batched over a leading chain axis and run on the GPU
(csrc/mlb_dense.cu through `mlb_gp_*` / `mlb_dense_cholesky` / the dense helpers of
`linalg`).  The closures keep the reference's names and argument order; the blocks are its
tuple `(dy_du [n, Y, U], dy_dn = chol_covar [n, Y, Y])`.

What is NOT here is a model: the reference's `covar_func` is an arbitrary Python closure
that cannot cross the C ABI, so `jacob_constr_blocks` takes the covariance matrix and its
derivatives in each latent direction (what `jax.jvp` of `covar_func` gives the reference,
systems.py:866-869) from the caller.  The gradient of `log_det_sqrt_gram` (reverse mode
through the Cholesky factorisation in the reference) is not built."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, stream_ptr
from .linalg import _dense, jvp_cholesky_mtx_mult_by_vct


def _p(t):
    return C.c_void_p(t.data_ptr())


def _blocks(dy_du, dy_dn):
    _lib.require_cuda()
    du = torch.as_tensor(dy_du, dtype=torch.float64)
    if du.ndim != 3:
        raise _lib.MlbError("dy_du must be [n_chain, dim_y, dim_u]")
    n, Y, U = du.shape
    return n, Y, U, _dense(du, (Y, U)), _dense(dy_dn, (Y, Y))


def cholesky(mtx):
    """Lower Cholesky factor per chain, mtx [n, dim, dim] symmetric positive definite (a
    matrix that is not gives NaNs for its chain)."""
    _lib.require_cuda()
    m = torch.as_tensor(mtx, dtype=torch.float64)
    n, dim = m.shape[0], m.shape[1]
    md = _dense(m, (dim, dim))
    out = torch.empty_like(md)
    check(lib().mlb_dense_cholesky(C.c_int64(n), C.c_int32(dim), _p(md), _p(out), stream_ptr()))
    return out


def jacob_constr_blocks(covar, dcovar_du, noise, y_obs):
    """systems.py:864-876 given covar [n, Y, Y] and dcovar_du [n, U, Y, Y] (the derivative of
    the covariance in every latent direction): ((dy_du, chol_covar), constr)."""
    chol_covar = cholesky(covar)
    n, Y = chol_covar.shape[0], chol_covar.shape[1]
    nz = _dense(noise, (Y,))
    dy_du = jvp_cholesky_mtx_mult_by_vct(dcovar_du, chol_covar, nz).permute(0, 2, 1).contiguous()
    # chol_covar @ n - y_obs through the Jacobian product with a zero latent part
    vq = torch.cat([torch.zeros(n, dy_du.shape[2], dtype=torch.float64, device="cuda"), nz], 1)
    c = lmult_by_jacob_constr(dy_du, chol_covar, vq) - torch.as_tensor(
        y_obs, dtype=torch.float64, device="cuda")
    return (dy_du, chol_covar), c


def _jacob_mult(dy_du, dy_dn, vct, transpose):
    n, Y, U, du, dn = _blocks(dy_du, dy_dn)
    v = _dense(vct, (Y,) if transpose else (U + Y,))
    out = torch.empty(n, U + Y if transpose else Y, dtype=torch.float64, device="cuda")
    check(lib().mlb_gp_jacob_constr_mult(
        C.c_int64(n), C.c_int32(Y), C.c_int32(U), C.c_int32(int(transpose)), _p(du), _p(dn),
        _p(v), _p(out), stream_ptr()))  # fmt: skip
    return out


def lmult_by_jacob_constr(dy_du, dy_dn, vct):  # systems.py:878-880
    return _jacob_mult(dy_du, dy_dn, vct, False)


def rmult_by_jacob_constr(dy_du, dy_dn, vct):  # systems.py:882-883
    return _jacob_mult(dy_du, dy_dn, vct, True)


def decompose_gram(dy_du, dy_dn):  # systems.py:885-888
    n, Y, U, du, dn = _blocks(dy_du, dy_dn)
    chol_cap = torch.empty(n, U, U, dtype=torch.float64, device="cuda")
    check(lib().mlb_gp_decompose_gram(C.c_int64(n), C.c_int32(Y), C.c_int32(U), _p(du), _p(dn),
                                      _p(chol_cap), stream_ptr()))
    return (chol_cap,)


def lmult_by_inv_gram(dy_du, dy_dn, chol_cap_mtx, vct):  # systems.py:890-897
    n, Y, U, du, dn = _blocks(dy_du, dy_dn)
    cc, v = _dense(chol_cap_mtx, (U, U)), _dense(vct, (Y,))
    out = torch.empty(n, Y, dtype=torch.float64, device="cuda")
    check(lib().mlb_gp_lmult_by_inv_gram(C.c_int64(n), C.c_int32(Y), C.c_int32(U), _p(du),
                                         _p(dn), _p(cc), _p(v), _p(out), stream_ptr()))
    return out


def lmult_by_inv_jacob_product(dy_du_l, dy_dn_l, dy_du_r, dy_dn_r, vct):  # systems.py:899-911
    n, Y, U, dul, dnl = _blocks(dy_du_l, dy_dn_l)
    dur, dnr, v = _dense(dy_du_r, (Y, U)), _dense(dy_dn_r, (Y, Y)), _dense(vct, (Y,))
    out = torch.empty(n, Y, dtype=torch.float64, device="cuda")
    check(lib().mlb_gp_lmult_by_inv_jacob_product(
        C.c_int64(n), C.c_int32(Y), C.c_int32(U), _p(dul), _p(dnl), _p(dur), _p(dnr), _p(v),
        _p(out), stream_ptr()))  # fmt: skip
    return out


def log_det_sqrt_gram(dy_du, dy_dn, chol_cap_mtx=None):
    """systems.py:913-918 from the blocks (the reference's closure takes q and calls
    jacob_constr_blocks itself)."""
    n, Y, U, du, dn = _blocks(dy_du, dy_dn)
    cc = decompose_gram(du, dn)[0] if chol_cap_mtx is None else _dense(chol_cap_mtx, (U, U))
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    check(lib().mlb_gp_log_det_sqrt_gram(C.c_int64(n), C.c_int32(Y), C.c_int32(U), _p(dn),
                                         _p(cc), _p(out), stream_ptr()))
    return out
