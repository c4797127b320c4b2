"""`mlift.linalg` (reference mlift/linalg.py), batched over a leading chain axis and run on
the GPU: the tridiagonal primitives (linalg.py:46-114, csrc/mlb_linalg.cu) under the
state-space systems' "tridiagonal + low rank" Gram matrices (mlift/systems.py:1189-1238) and
the dense helpers (linalg.py:6-43, csrc/mlb_dense.cu) under the Gaussian-process systems
(systems.py:871-899).  Same names and argument order as the reference; every array carries
a leading chain axis.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, p_soa, stream_ptr


def _rows(x, n):
    """[n, rows] (or [rows] broadcast over chains) -> component-major device tensor."""
    t = torch.as_tensor(x, dtype=torch.float64)
    if t.ndim == 1:
        t = t[None].expand(n, -1)
    return _lib.to_soa(t)


def tridiagonal_solve(a, b, c, d):
    """Solve T x = d per chain, T tridiagonal with lower / main / upper diagonals a
    [n, dim-1], b [n, dim], c [n, dim-1]; d [n, dim] or [n, dim, n_rhs] (the reference
    vmaps over right-hand sides).  linalg.py:46-85."""
    b_t = torch.as_tensor(b, dtype=torch.float64)
    n, dim = b_t.shape
    d_t = torch.as_tensor(d, dtype=torch.float64)
    multi = d_t.ndim == 3
    n_rhs = d_t.shape[2] if multi else 1
    # rows r * dim + k of one component-major array
    d_rows = (d_t.permute(0, 2, 1).reshape(n, n_rhs * dim) if multi else d_t)
    bs, ds = _rows(b_t, n), _rows(d_rows, n)
    if dim > 1:
        as_, cs = _rows(a, n), _rows(c, n)
        pa, lda = p_soa(as_, dim - 1)
        pc, ldc = p_soa(cs, dim - 1)
    else:
        pa = pc = None
        lda = ldc = None
    pb, ld = p_soa(bs, dim)
    pd, ldd = p_soa(ds, n_rhs * dim)
    x = _lib.soa_empty(n, n_rhs * dim, "cuda")
    px, ldx = p_soa(x, n_rhs * dim)
    if len({v for v in (ld, ldd, ldx, lda, ldc) if v is not None}) != 1:
        raise _lib.MlbError("tridiagonal_solve: arrays must share their leading dimension")
    check(lib().mlb_tridiagonal_solve(C.c_int64(n), C.c_int32(dim), C.c_int32(n_rhs),
                                      C.c_int64(ld), pa, pb, pc, pd, px, stream_ptr()))
    return x.reshape(n, n_rhs, dim).permute(0, 2, 1) if multi else x


def tridiagonal_pos_def_log_det(a, b):
    """log det of the symmetric positive-definite tridiagonal matrix with off-diagonal a
    [n, dim-1] and main diagonal b [n, dim] (all positive), per chain.  linalg.py:88-114."""
    b_t = torch.as_tensor(b, dtype=torch.float64)
    n, dim = b_t.shape
    bs = _rows(b_t, n)
    pb, ld = p_soa(bs, dim)
    pa = None
    if dim > 1:
        as_ = _rows(a, n)
        pa, lda = p_soa(as_, dim - 1)
        if lda != ld:
            raise _lib.MlbError("tridiagonal_pos_def_log_det: a and b must share ld")
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    check(lib().mlb_tridiagonal_pos_def_log_det(C.c_int64(n), C.c_int32(dim), C.c_int64(ld), pa,
                                                pb, C.c_void_p(out.data_ptr()), stream_ptr()))
    return out


def _dense(x, shape_tail):
    t = torch.as_tensor(x, dtype=torch.float64)
    if t.shape[1:] != tuple(shape_tail):
        raise _lib.MlbError(f"expected trailing shape {tuple(shape_tail)}, got {tuple(t.shape[1:])}")
    return t.to("cuda").contiguous()


def lu_triangular_solve(l, u, b):
    """Solve `l @ u @ x = b` per chain, `l` lower- and `u` upper-triangular [n, dim, dim];
    b [n, dim] or [n, dim, n_rhs].  linalg.py:6-9."""
    _lib.require_cuda()
    l_t = torch.as_tensor(l, dtype=torch.float64)
    n, dim = l_t.shape[0], l_t.shape[1]
    b_t = torch.as_tensor(b, dtype=torch.float64)
    vec = b_t.ndim == 2
    n_rhs = 1 if vec else b_t.shape[2]
    ld, ud = _dense(l_t, (dim, dim)), _dense(u, (dim, dim))
    bd = _dense(b_t.reshape(n, dim, n_rhs), (dim, n_rhs))
    if ud.shape[0] != n or bd.shape[0] != n:
        raise _lib.MlbError("lu_triangular_solve: arrays must share the chain axis")
    x = torch.empty_like(bd)
    check(lib().mlb_lu_triangular_solve(
        C.c_int64(n), C.c_int32(dim), C.c_int32(n_rhs), C.c_void_p(ld.data_ptr()),
        C.c_void_p(ud.data_ptr()), C.c_void_p(bd.data_ptr()), C.c_void_p(x.data_ptr()),
        stream_ptr()))  # fmt: skip
    return x[:, :, 0] if vec else x


def jvp_cholesky_mtx_mult_by_vct(tangents, chol_mtx, vct):
    """Forward-mode derivative of `mtx -> cholesky(mtx) @ vct` in the symmetric direction
    `tangents` [n, dim, dim] - or [n, n_dir, dim, dim] for several directions at once, the
    reference's `vmap(..., (1, None, None), 1)` of systems.py:871-873 - given
    `chol_mtx = cholesky(mtx)` [n, dim, dim] and vct [n, dim].  linalg.py:12-43."""
    _lib.require_cuda()
    c_t = torch.as_tensor(chol_mtx, dtype=torch.float64)
    n, dim = c_t.shape[0], c_t.shape[1]
    t_t = torch.as_tensor(tangents, dtype=torch.float64)
    single = t_t.ndim == 3
    n_dir = 1 if single else t_t.shape[1]
    td = _dense(t_t.reshape(n, n_dir, dim, dim), (n_dir, dim, dim))
    cd, vd = _dense(c_t, (dim, dim)), _dense(vct, (dim,))
    if td.shape[0] != n or vd.shape[0] != n:
        raise _lib.MlbError("jvp_cholesky_mtx_mult_by_vct: arrays must share the chain axis")
    out = torch.empty(n, n_dir, dim, dtype=torch.float64, device="cuda")
    check(lib().mlb_jvp_cholesky_mtx_mult_by_vct(
        C.c_int64(n), C.c_int32(n_dir), C.c_int32(dim), C.c_void_p(td.data_ptr()),
        C.c_void_p(cd.data_ptr()), C.c_void_p(vd.data_ptr()), C.c_void_p(out.data_ptr()),
        stream_ptr()))  # fmt: skip
    return out[:, 0] if single else out
