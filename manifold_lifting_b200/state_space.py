""""Tridiagonal + low rank" Gram algebra of the reference's state-space systems (the
closures of `PartiallyInvertibleStateSpaceModelSystem`, mlift/systems.py:1161-1243), batched
over a leading chain axis and run on the GPU (csrc/mlb_ssm.cu through the C ABI,
include/mlb.h `mlb_ssm_*`).  Same function names and argument order as the reference's
closures; every array carries a leading chain axis:

    dc_du [n, dim_y, dim_u], dc_dv [n, dim_y], dc_dn = (dc_dn0 [n, dim_y], dc_dn1 [n, dim_y-1]),
    dx_dy [n, dim_y]  (the tuple `jacob_constr_split_blocks` returns, systems.py:1136-1150)

`JacobBlocks` converts such a tuple once to the component-major device layout; the functions
take either the raw arrays or a `JacobBlocks` in place of the four block arguments' first.
The model functions that produce the blocks (`constr_split`, a Python closure in the
reference) are not part of this module.  There is no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, p_soa, stream_ptr


class _CBlocks(C.Structure):
    _fields_ = [("dc_du", C.c_void_p), ("dc_dv", C.c_void_p), ("dc_dn0", C.c_void_p),
                ("dc_dn1", C.c_void_p), ("dx_dy", C.c_void_p)]  # fmt: skip


class JacobBlocks:
    """Device copy, in the component-major layout, of one batch of Jacobian blocks."""

    def __init__(self, dc_du, dc_dv, dc_dn, dx_dy=None):
        _lib.require_cuda()
        du = torch.as_tensor(dc_du, dtype=torch.float64)
        if du.ndim != 3:
            raise _lib.MlbError("dc_du must be [n_chain, dim_y, dim_u]")
        self.n, self.dim_y, self.dim_u = du.shape
        self.dc_du = _lib.to_soa(du.reshape(self.n, self.dim_y * self.dim_u))
        self.dc_dv = self._rows(dc_dv, self.dim_y)
        self.dc_dn0 = self._rows(dc_dn[0], self.dim_y)
        self.dc_dn1 = self._rows(dc_dn[1], self.dim_y - 1) if self.dim_y > 1 else None
        self.dx_dy = self._rows(dx_dy, self.dim_y) if dx_dy is not None else None

    def _rows(self, x, rows):
        t = _lib.to_soa(torch.as_tensor(x, dtype=torch.float64))
        if tuple(t.shape) != (self.n, rows):
            raise _lib.MlbError(f"expected a [{self.n}, {rows}] block, got {tuple(t.shape)}")
        return t

    def c_struct(self):
        def ptr(t, rows):
            if t is None:
                return None
            p, ld = p_soa(t, rows)
            if ld != self.ld:
                raise _lib.MlbError("state-space blocks must share their leading dimension")
            return p

        return _CBlocks(ptr(self.dc_du, self.dim_y * self.dim_u), ptr(self.dc_dv, self.dim_y),
                        ptr(self.dc_dn0, self.dim_y), ptr(self.dc_dn1, self.dim_y - 1),
                        ptr(self.dx_dy, self.dim_y))  # fmt: skip

    @property
    def ld(self):
        return p_soa(self.dc_dv, self.dim_y)[1]


def _blocks(dc_du, dc_dv=None, dc_dn=None, dx_dy=None):
    return dc_du if isinstance(dc_du, JacobBlocks) else JacobBlocks(dc_du, dc_dv, dc_dn, dx_dy)


def _dims(J):
    return C.c_int64(J.n), C.c_int32(J.dim_y), C.c_int32(J.dim_u), C.c_int64(J.ld)


def _vec(J, x, rows):
    t = _lib.to_soa(torch.as_tensor(x, dtype=torch.float64))
    if tuple(t.shape) != (J.n, rows):
        raise _lib.MlbError(f"expected a [{J.n}, {rows}] array, got {tuple(t.shape)}")
    p, ld = p_soa(t, rows)
    if ld != J.ld:
        raise _lib.MlbError("arrays must share the blocks' leading dimension")
    return t, p


def lmult_by_jacob_constr(dc_du, dc_dv, dc_dn, dx_dy, vct):
    """J vct per chain, vct [n, dim_u + 2 dim_y].  systems.py:1161-1175."""
    J = _blocks(dc_du, dc_dv, dc_dn, dx_dy)
    _, pv = _keep = _vec(J, vct, J.dim_u + 2 * J.dim_y)
    out = _lib.soa_empty(J.n, J.dim_y)
    cs = J.c_struct()
    check(lib().mlb_ssm_lmult_by_jacob_constr(*_dims(J), C.byref(cs), pv, p_soa(out)[0],
                                              stream_ptr()))  # fmt: skip
    return out


def rmult_by_jacob_constr(dc_du, dc_dv, dc_dn, dx_dy, vct):
    """J^T vct per chain, vct [n, dim_y].  systems.py:1177-1187."""
    J = _blocks(dc_du, dc_dv, dc_dn, dx_dy)
    _, pv = _keep = _vec(J, vct, J.dim_y)
    out = _lib.soa_empty(J.n, J.dim_u + 2 * J.dim_y)
    cs = J.c_struct()
    check(lib().mlb_ssm_rmult_by_jacob_constr(*_dims(J), C.byref(cs), pv, p_soa(out)[0],
                                              stream_ptr()))  # fmt: skip
    return out


def decompose_gram(dc_du, dc_dv=None, dc_dn=None, dx_dy=None):
    """(a [n, dim_y-1], b [n, dim_y], chol_cap_mtx [n, dim_u, dim_u]).  systems.py:1189-1197."""
    J = _blocks(dc_du, dc_dv, dc_dn, dx_dy)
    a = _lib.soa_empty(J.n, max(J.dim_y - 1, 1))
    b = _lib.soa_empty(J.n, J.dim_y)
    chol = _lib.soa_empty(J.n, J.dim_u * J.dim_u)
    cs = J.c_struct()
    check(lib().mlb_ssm_decompose_gram(*_dims(J), C.byref(cs), p_soa(a)[0], p_soa(b)[0],
                                       p_soa(chol)[0], stream_ptr()))  # fmt: skip
    return a[:, : J.dim_y - 1], b, chol.reshape(J.n, J.dim_u, J.dim_u)


def _gram_ptrs(J, a, b, chol_cap_mtx):
    b_t, pb = _vec(J, b, J.dim_y)
    keep = [b_t]
    pa = None
    if J.dim_y > 1:
        a_t, pa = _vec(J, a, J.dim_y - 1)
        keep.append(a_t)
    ch = torch.as_tensor(chol_cap_mtx, dtype=torch.float64).reshape(J.n, J.dim_u * J.dim_u)
    ch_t, pc = _vec(J, ch, J.dim_u * J.dim_u)
    keep.append(ch_t)
    return keep, pa, pb, pc


def lmult_by_inv_gram(dc_du, dc_dv, dc_dn, dx_dy, a, b, chol_cap_mtx, vct):
    """(J J^T)^{-1} vct per chain.  systems.py:1199-1210."""
    J = _blocks(dc_du, dc_dv, dc_dn, dx_dy)
    _keep, pa, pb, pc = _gram_ptrs(J, a, b, chol_cap_mtx)
    _, pv = _keep2 = _vec(J, vct, J.dim_y)
    out = _lib.soa_empty(J.n, J.dim_y)
    cs = J.c_struct()
    check(lib().mlb_ssm_lmult_by_inv_gram(*_dims(J), C.byref(cs), pa, pb, pc, pv,
                                          p_soa(out)[0], stream_ptr()))  # fmt: skip
    return out


def lmult_by_inv_jacob_product(dc_du_l, dc_dv_l, dc_dn_l, dx_dy_l, dc_du_r, dc_dv_r, dc_dn_r,
                               dx_dy_r, vct):
    """(J_l J_r^T)^{-1} vct per chain.  systems.py:1212-1232."""
    Jl = _blocks(dc_du_l, dc_dv_l, dc_dn_l, dx_dy_l)
    Jr = _blocks(dc_du_r, dc_dv_r, dc_dn_r, dx_dy_r)
    if (Jl.n, Jl.dim_y, Jl.dim_u, Jl.ld) != (Jr.n, Jr.dim_y, Jr.dim_u, Jr.ld):
        raise _lib.MlbError("left and right blocks differ in shape")
    _, pv = _keep = _vec(Jl, vct, Jl.dim_y)
    out = _lib.soa_empty(Jl.n, Jl.dim_y)
    cl, cr = Jl.c_struct(), Jr.c_struct()
    check(lib().mlb_ssm_lmult_by_inv_jacob_product(*_dims(Jl), C.byref(cl), C.byref(cr), pv,
                                                   p_soa(out)[0], stream_ptr()))  # fmt: skip
    return out


def log_det_sqrt_gram(dc_du, dc_dv=None, dc_dn=None, dx_dy=None, gram=None):
    """sum log diag(chol_cap) + tridiagonal_pos_def_log_det(a, b) / 2 - sum log|dx_dy| per
    chain, from the blocks (the reference's closure evaluates them from q first,
    systems.py:1234-1243); `gram` = a decomposition already at hand."""
    J = _blocks(dc_du, dc_dv, dc_dn, dx_dy)
    if J.dx_dy is None:
        raise _lib.MlbError("log_det_sqrt_gram needs dx_dy")
    a, b, chol = gram if gram is not None else decompose_gram(J)
    _keep, pa, pb, pc = _gram_ptrs(J, a, b, chol)
    out = torch.empty(J.n, dtype=torch.float64, device="cuda")
    cs = J.c_struct()
    check(lib().mlb_ssm_log_det_sqrt_gram(*_dims(J), C.byref(cs), pa, pb, pc,
                                          C.c_void_p(out.data_ptr()), stream_ptr()))  # fmt: skip
    return out
