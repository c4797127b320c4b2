"""NumPy restatement of the tridiagonal primitives of mlift/linalg.py:46-114 (and of
mlift/math.py:7-19, which they use), one system at a time, with the reference's scans
written as loops in the same order.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def log1m_exp(val):  # mlift/math.py:7-14
    return np.log(-np.expm1(val)) if val > np.log(2.0) else np.log1p(-np.exp(val))


def log_diff_exp(val1, val2):  # mlift/math.py:17-19
    return val1 + log1m_exp(val2 - val1)


def tridiagonal_solve(a, b, c, d):
    """linalg.py:46-85: forward scan (:67-75), backward scan (:80-85)."""
    a, b, c, d = (np.asarray(v, dtype=np.float64) for v in (a, b, c, d))
    dim = b.shape[0]
    b_, d_ = b.copy(), d.copy()
    for i in range(1, dim):
        w = a[i - 1] / b_[i - 1]
        b_[i] = b[i] - w * c[i - 1]
        d_[i] = d[i] - w * d_[i - 1]
    x = np.empty_like(d_)
    x[-1] = d_[-1] / b_[-1]
    for i in range(dim - 2, -1, -1):
        x[i] = (d_[i] - c[i] * x[i + 1]) / b_[i]
    return x


def tridiagonal_pos_def_log_det(a, b):
    """linalg.py:88-114: log-continuant recursion."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    l_i, l_im1 = np.log(b[0]), 0.0
    for i in range(b.shape[0] - 1):
        l_ip1 = log_diff_exp(np.log(b[i + 1]) + l_i, 2 * np.log(abs(a[i])) + l_im1)
        l_i, l_im1 = l_ip1, l_i
    return l_i


def lu_triangular_solve(l, u, b):
    """linalg.py:6-9: solve l u x = b, l lower- and u upper-triangular (two substitutions)."""
    from scipy.linalg import solve_triangular

    x = solve_triangular(np.asarray(l, dtype=np.float64), np.asarray(b, dtype=np.float64), lower=True)
    return solve_triangular(np.asarray(u, dtype=np.float64), x, lower=False)


def jvp_cholesky_mtx_mult_by_vct(tangents, chol_mtx, vct):
    """linalg.py:12-43: chol Phi(chol^-1 t chol^-T) vct with Phi = lower triangle, diagonal
    halved (forward-mode rule of the Cholesky factorisation)."""
    from scipy.linalg import solve_triangular

    L = np.asarray(chol_mtx, dtype=np.float64)
    tmp = solve_triangular(L, np.asarray(tangents, dtype=np.float64).T, lower=True).T  # t L^-T
    tmp = solve_triangular(L, tmp, lower=True)
    phi = np.tril(tmp) / (np.ones_like(L) + np.identity(L.shape[-1]))
    return L @ (phi @ np.asarray(vct, dtype=np.float64))
