"""Dense helpers of mlift.linalg (lu_triangular_solve, jvp_cholesky_mtx_mult_by_vct,
linalg.py:6-43) and the Gaussian-process block algebra built on them (systems.py:859-918):
the NumPy restatements of oracle/mlift_oracle against the outputs of the REFERENCE'S OWN
source (tests/golden/ref_dense.npz, written by tests/golden/make_ref_golden.py dense, which
runs /root/reference/mlift on the jax stand-in)."""
import os

import numpy as np

from mlift_oracle import gp as ogp
from mlift_oracle import linalg as ol

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dense.npz"))


def _rel(x, ref):
    return np.abs(np.asarray(x) - ref).max() / max(1.0, np.abs(ref).max())


def test_dense_helper_restatements_match_reference_source():
    for k in range(int(G["n_case"])):
        l, u, b = G[f"l_{k}"], G[f"u_{k}"], G[f"b_{k}"]
        assert _rel(ol.lu_triangular_solve(l, u, b), G[f"x_{k}"]) < 1e-12
        assert _rel(ol.lu_triangular_solve(l, u, b[:, 0]), G[f"xv_{k}"]) < 1e-12
        # and it is a solve: l u x = b
        assert _rel(l @ (u @ G[f"x_{k}"]), b) < 1e-10
        t, v = G[f"t_{k}"], G[f"v_{k}"]
        jvp = np.stack([ol.jvp_cholesky_mtx_mult_by_vct(t[:, d], l, v)
                        for d in range(t.shape[1])], 1)  # fmt: skip
        assert _rel(jvp, G[f"jvp_{k}"]) < 1e-12
        # against central differences of mtx -> cholesky(mtx) v
        mtx, eps = l @ l.T, 1e-6
        fd = (np.linalg.cholesky(mtx + eps * t[:, 0]) - np.linalg.cholesky(mtx - eps * t[:, 0])) @ v / (2 * eps)
        assert _rel(jvp[:, 0], fd) < 1e-6


def test_gaussian_process_closures_match_reference_source():
    y_obs = G["gp_y_obs"]
    for i in range(G["gp_q"].shape[0]):
        q, vq, wy = G["gp_q"][i], G["gp_vec_q"][i], G["gp_vec_y"][i]
        dim_u = G["gp_dy_du"].shape[2]
        (dy_du, dy_dn), c = ogp.jacob_constr_blocks(
            G["gp_covar"][i], G["gp_dcovar_du"][i], q[dim_u:], y_obs)
        assert _rel(dy_du, G["gp_dy_du"][i]) < 1e-11 and _rel(dy_dn, G["gp_dy_dn"][i]) < 1e-12
        assert _rel(c, G["gp_c"][i]) < 1e-12
        (chol_cap,) = ogp.decompose_gram(dy_du, dy_dn)
        assert _rel(chol_cap, G["gp_chol_cap"][i]) < 1e-11
        assert abs(ogp.log_det_sqrt_gram(dy_du, dy_dn) - G["gp_logdet"][i]) < 1e-11
        assert _rel(ogp.lmult_by_jacob_constr(dy_du, dy_dn, vq), G["gp_jv"][i]) < 1e-12
        assert _rel(ogp.rmult_by_jacob_constr(dy_du, dy_dn, wy), G["gp_jtw"][i]) < 1e-12
        pinv = ogp.rmult_by_jacob_constr(dy_du, dy_dn, ogp.lmult_by_inv_gram(dy_du, dy_dn, chol_cap, wy))
        assert _rel(pinv, G["gp_pinv_w"][i]) < 1e-10
        ijp = ogp.lmult_by_inv_jacob_product(dy_du, dy_dn, G["gp_dy_du_prev"][i],
                                             G["gp_dy_dn_prev"][i], wy)  # fmt: skip
        assert _rel(ijp, G["gp_ijp_w"][i]) < 1e-10
        nsc = ogp.rmult_by_jacob_constr(dy_du, dy_dn, ogp.lmult_by_inv_gram(
            dy_du, dy_dn, chol_cap, ogp.lmult_by_jacob_constr(dy_du, dy_dn, vq)))
        assert _rel(vq - nsc, G["gp_proj_v"][i]) < 1e-10
