"""NumPy restatement, one chain at a time, of the "tridiagonal + low rank" Gram algebra of
the reference's state-space systems (mlift/systems.py:1161-1243, the closures of
PartiallyInvertibleStateSpaceModelSystem), on top of the restated tridiagonal scans
(linalg.py here, mlift/linalg.py:46-114).  Blocks are the reference's tuple
(dc_du [Y, U], dc_dv [Y], (dc_dn0 [Y], dc_dn1 [Y - 1]), dx_dy [Y]).  `dense_jacobian`
assembles the Y x Q Jacobian those blocks stand for (systems.py:1161-1187 read as a
matrix), so the tests can hold every function to dense linear algebra.
TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.linalg as sla

from .linalg import tridiagonal_pos_def_log_det, tridiagonal_solve


def lmult_by_jacob_constr(dc_du, dc_dv, dc_dn, dx_dy, vct):  # systems.py:1161-1175
    dim_y, dim_u = dc_du.shape
    vct_u, vct_v, vct_n = vct[:dim_u], vct[dim_u : dim_u + dim_y], vct[dim_u + dim_y :]
    return (dc_du @ vct_u + dc_dv * vct_v + dc_dn[0] * vct_n
            + np.pad(dc_dn[1] * vct_n[:-1], (1, 0)))  # fmt: skip


def rmult_by_jacob_constr(dc_du, dc_dv, dc_dn, dx_dy, vct):  # systems.py:1177-1187
    return np.concatenate(
        (vct @ dc_du, vct * dc_dv, vct * dc_dn[0] + np.pad(dc_dn[1] * vct[1:], (0, 1)))
    )


def decompose_gram(dc_du, dc_dv, dc_dn, dx_dy):  # systems.py:1189-1197
    dim_u = dc_du.shape[1]
    a = dc_dn[0][:-1] * dc_dn[1]
    b = dc_dv**2 + dc_dn[0] ** 2 + np.pad(dc_dn[1] ** 2, (1, 0))
    sol = np.stack([tridiagonal_solve(a, b, a, dc_du[:, j]) for j in range(dim_u)], 1)
    cap_mtx = np.eye(dim_u) + dc_du.T @ sol
    return a, b, np.linalg.cholesky(cap_mtx)


def lmult_by_inv_gram(dc_du, dc_dv, dc_dn, dx_dy, a, b, chol_cap_mtx, vct):  # :1199-1210
    return tridiagonal_solve(
        a, b, a,
        vct - dc_du @ sla.cho_solve((chol_cap_mtx, True), dc_du.T @ tridiagonal_solve(a, b, a, vct)),
    )  # fmt: skip


def lmult_by_inv_jacob_product(dc_du_l, dc_dv_l, dc_dn_l, dx_dy_l, dc_du_r, dc_dv_r, dc_dn_r,
                               dx_dy_r, vct):  # systems.py:1212-1232
    dim_u = dc_du_l.shape[1]
    a = dc_dn_l[1] * dc_dn_r[0][:-1]
    b = dc_dv_l * dc_dv_r + dc_dn_l[0] * dc_dn_r[0] + np.pad(dc_dn_l[1] * dc_dn_r[1], (1, 0))
    c = dc_dn_l[0][:-1] * dc_dn_r[1]
    sol = np.stack([tridiagonal_solve(a, b, c, dc_du_l[:, j]) for j in range(dim_u)], 1)
    cap_mtx = np.eye(dim_u) + dc_du_r.T @ sol
    return tridiagonal_solve(
        a, b, c,
        vct - dc_du_l @ np.linalg.solve(cap_mtx, dc_du_r.T @ tridiagonal_solve(a, b, c, vct)),
    )  # fmt: skip


def log_det_sqrt_gram(dc_du, dc_dv, dc_dn, dx_dy):  # systems.py:1234-1243 (from the blocks)
    a, b, chol_cap_mtx = decompose_gram(dc_du, dc_dv, dc_dn, dx_dy)
    return (np.log(chol_cap_mtx.diagonal()).sum() + tridiagonal_pos_def_log_det(a, b) / 2
            - np.log(abs(dx_dy)).sum())  # fmt: skip


def dense_jacobian(dc_du, dc_dv, dc_dn, dx_dy):
    """[Y, U + 2Y] matrix whose products are lmult / rmult_by_jacob_constr."""
    dim_y = dc_du.shape[0]
    dn = np.diag(dc_dn[0])
    if dim_y > 1:
        dn = dn + np.diag(dc_dn[1], -1)
    return np.concatenate((dc_du, np.diag(dc_dv), dn), 1)


def random_blocks(rng, dim_y, dim_u):
    """Blocks shaped like a state-space model's (|dc_dv|, |dc_dn0| of order one, the
    sub-diagonal coupling smaller): the tridiagonal part is positive definite with positive
    off-diagonals, as the reference's log-continuant recursion assumes (linalg.py:88-103)."""
    dc_du = rng.normal(size=(dim_y, dim_u))
    dc_dv = 0.5 + np.abs(rng.normal(size=dim_y))
    dn0 = 0.5 + np.abs(rng.normal(size=dim_y))
    dn1 = 0.1 + 0.4 * np.abs(rng.normal(size=max(dim_y - 1, 0)))
    dx_dy = rng.normal(size=dim_y) + np.sign(rng.normal(size=dim_y)) * 0.3
    return dc_du, dc_dv, (dn0, dn1), dx_dy
