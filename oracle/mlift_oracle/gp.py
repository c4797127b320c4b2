"""NumPy restatement of the closures of `GaussianProcessModelSystem`
(mlift/systems.py:859-918: y = cholesky(covar(u)) n; Jacobian blocks (dy_du [Y, U], dy_dn =
chol_covar [Y, Y])), one chain at a time, given the blocks - plus `jacob_constr_blocks` for a
covariance whose directional derivatives are supplied (systems.py:864-876).
TEST INFRASTRUCTURE ONLY."""
import numpy as np
from scipy.linalg import cho_solve, solve_triangular

from .linalg import jvp_cholesky_mtx_mult_by_vct, lu_triangular_solve


def jacob_constr_blocks(covar, dcovar_du, n, y_obs):  # systems.py:864-876
    """dcovar_du [Y, U, Y]: derivative of the covariance in each latent direction."""
    chol_covar = np.linalg.cholesky(covar)
    dy_du = np.stack([jvp_cholesky_mtx_mult_by_vct(dcovar_du[:, k], chol_covar, n)
                      for k in range(dcovar_du.shape[1])], 1)  # fmt: skip
    return (dy_du, chol_covar), chol_covar @ n - y_obs


def lmult_by_jacob_constr(dy_du, dy_dn, vct):  # systems.py:878-880
    dim_u = dy_du.shape[1]
    return dy_du @ vct[:dim_u] + dy_dn @ vct[dim_u:]


def rmult_by_jacob_constr(dy_du, dy_dn, vct):  # systems.py:882-883
    return np.concatenate((vct @ dy_du, vct @ dy_dn))


def decompose_gram(dy_du, dy_dn):  # systems.py:885-888
    cap_mtx = np.eye(dy_du.shape[1]) + dy_du.T @ cho_solve((dy_dn, True), dy_du)
    return (np.linalg.cholesky(cap_mtx),)


def lmult_by_inv_gram(dy_du, dy_dn, chol_cap_mtx, vct):  # systems.py:890-897
    return cho_solve(
        (dy_dn, True),
        vct - dy_du @ cho_solve((chol_cap_mtx, True), dy_du.T @ cho_solve((dy_dn, True), vct)))


def lmult_by_inv_jacob_product(dy_du_l, dy_dn_l, dy_du_r, dy_dn_r, vct):  # systems.py:899-911
    cap_mtx = np.eye(dy_du_l.shape[1]) + dy_du_r.T @ lu_triangular_solve(dy_dn_l, dy_dn_r.T, dy_du_l)
    return lu_triangular_solve(
        dy_dn_l, dy_dn_r.T,
        vct - dy_du_l @ np.linalg.solve(
            cap_mtx, dy_du_r.T @ lu_triangular_solve(dy_dn_l, dy_dn_r.T, vct)))


def log_det_sqrt_gram(dy_du, dy_dn):  # systems.py:913-918 (from the blocks)
    (chol_cap_mtx,) = decompose_gram(dy_du, dy_dn)
    return np.log(chol_cap_mtx.diagonal()).sum() + np.log(dy_dn.diagonal()).sum()
