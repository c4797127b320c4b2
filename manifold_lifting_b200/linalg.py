"""`mlift.linalg` tridiagonal primitives (reference mlift/linalg.py:46-114), batched over a
leading chain axis and run on the GPU (csrc/mlb_linalg.cu): the linear algebra under the
state-space systems' "tridiagonal + low rank" Gram matrices (mlift/systems.py:1189-1238).
Same names and argument order as the reference; every array carries a leading chain axis.
`lu_triangular_solve` and `jvp_cholesky_mtx_mult_by_vct` (dense, used by the Gaussian-process
systems only) are not built.
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
