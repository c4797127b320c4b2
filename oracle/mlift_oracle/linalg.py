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
