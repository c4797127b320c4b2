"""Dense helpers of mlift.linalg on the GPU (csrc/mlb_dense.cu through the C ABI):
`lu_triangular_solve`, `jvp_cholesky_mtx_mult_by_vct` (linalg.py:6-43) against the outputs of
the reference's own source (tests/golden/ref_dense.npz) and against the NumPy restatement on
batches, including the Jacobian block dy_du of the reference's Gaussian-process system
(systems.py:864-876) rebuilt from the helper."""
import os

import numpy as np
import pytest
import torch

from mlift_oracle import linalg as ol

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dense.npz"))


def _rel(x, ref):
    x = x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    return np.abs(x - ref).max() / max(1.0, np.abs(ref).max())


def test_dense_helpers_match_reference_source():
    import manifold_lifting_b200 as mlb

    la = mlb.linalg
    for k in range(int(G["n_case"])):
        l, u, b = G[f"l_{k}"], G[f"u_{k}"], G[f"b_{k}"]
        # three copies as a batch: chains are independent
        x = la.lu_triangular_solve(np.stack([l] * 3), np.stack([u] * 3), np.stack([b] * 3))
        assert _rel(x[1], G[f"x_{k}"]) < 1e-12 and torch.equal(x[0], x[2])
        xv = la.lu_triangular_solve(l[None], u[None], b[None, :, 0])
        assert _rel(xv[0], G[f"xv_{k}"]) < 1e-12
        t, v = G[f"t_{k}"], G[f"v_{k}"]  # t [dim, n_dir, dim]
        jvp = la.jvp_cholesky_mtx_mult_by_vct(t.transpose(1, 0, 2)[None], l[None], v[None])
        assert _rel(jvp[0].T, G[f"jvp_{k}"]) < 1e-12
        one = la.jvp_cholesky_mtx_mult_by_vct(t[:, 0][None], l[None], v[None])
        assert _rel(one[0], G[f"jvp_{k}"][:, 0]) < 1e-12


@pytest.mark.parametrize("dim,n_rhs", [(1, 1), (3, 2), (31, 9), (64, 1), (129, 5), (300, 17)])
def test_dense_helpers_match_restatement_on_batches(dim, n_rhs):
    import manifold_lifting_b200 as mlb

    la = mlb.linalg
    rng = np.random.default_rng(100 * dim + n_rhs)
    n = 7
    a, c = rng.standard_normal((2, n, dim, dim))
    l = np.linalg.cholesky(a @ a.transpose(0, 2, 1) / dim + np.eye(dim))
    u = np.linalg.cholesky(c @ c.transpose(0, 2, 1) / dim + np.eye(dim)).transpose(0, 2, 1)
    # the triangle that is not read may hold anything
    l_dirty = l + np.triu(rng.standard_normal((dim, dim)), 1)[None]
    u_dirty = u + np.tril(rng.standard_normal((dim, dim)), -1)[None]
    b = rng.standard_normal((n, dim, n_rhs))
    x = la.lu_triangular_solve(l_dirty, u_dirty, b)
    ref = np.stack([ol.lu_triangular_solve(l[i], u[i], b[i]) for i in range(n)])
    assert _rel(x, ref) < 1e-11
    t = rng.standard_normal((n, 2, dim, dim))
    t = t + t.transpose(0, 1, 3, 2)
    v = rng.standard_normal((n, dim))
    jvp = la.jvp_cholesky_mtx_mult_by_vct(t, l_dirty, v)
    ref = np.stack([[ol.jvp_cholesky_mtx_mult_by_vct(t[i, d], l[i], v[i]) for d in range(2)]
                    for i in range(n)])  # fmt: skip
    assert _rel(jvp, ref) < 1e-11


def test_gaussian_process_jacobian_block_from_the_helper():
    """dy_du of `GaussianProcessModelSystem.jacob_constr_blocks` (systems.py:864-876) =
    jvp_cholesky_mtx_mult_by_vct vmapped over the directional derivatives of the covariance,
    for all four recorded states in one call."""
    import manifold_lifting_b200 as mlb

    dim_u = G["gp_dy_du"].shape[2]
    chol = G["gp_dy_dn"]
    dcov = G["gp_dcovar_du"].transpose(0, 2, 1, 3)  # [n, U, Y, Y]
    noise = G["gp_q"][:, dim_u:]
    dy_du = mlb.linalg.jvp_cholesky_mtx_mult_by_vct(dcov, chol, noise)  # [n, U, Y]
    assert _rel(dy_du.permute(0, 2, 1), G["gp_dy_du"]) < 1e-11


def test_dense_helpers_argument_checks():
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    with pytest.raises(_lib.MlbError):
        mlb.linalg.lu_triangular_solve(np.eye(3)[None], np.eye(4)[None], np.ones((1, 3)))
    with pytest.raises(_lib.MlbError):
        mlb.linalg.jvp_cholesky_mtx_mult_by_vct(np.eye(3)[None], np.eye(3)[None], np.ones((2, 3)))
    # a zero pivot poisons that chain only
    l = np.stack([np.eye(4), np.eye(4)])
    l[1, 2, 2] = 0.0
    x = mlb.linalg.lu_triangular_solve(l, np.stack([np.eye(4)] * 2), np.ones((2, 4)))
    assert torch.isfinite(x[0]).all() and not torch.isfinite(x[1]).all()


# --------------------------------------------- Gaussian-process block algebra (mlb_gp_*)
def test_gaussian_process_closures_match_reference_source():
    """`manifold_lifting_b200.gaussian_process` (the closures of GaussianProcessModelSystem,
    systems.py:859-918, on the GPU) against the outputs of the reference's own source for
    the squared-exponential regression model (ref_dense.npz gp_*), all recorded states as
    one batch."""
    import manifold_lifting_b200 as mlb

    gp = mlb.gaussian_process
    dim_u = G["gp_dy_du"].shape[2]
    q, vq, wy = G["gp_q"], G["gp_vec_q"], G["gp_vec_y"]
    dcov = G["gp_dcovar_du"].transpose(0, 2, 1, 3)  # [n, U, Y, Y]
    (dy_du, dy_dn), c = gp.jacob_constr_blocks(G["gp_covar"], dcov, q[:, dim_u:], G["gp_y_obs"])
    assert _rel(dy_dn, G["gp_dy_dn"]) < 1e-12 and _rel(dy_du, G["gp_dy_du"]) < 1e-11
    assert _rel(c, G["gp_c"]) < 1e-12
    (chol_cap,) = gp.decompose_gram(dy_du, dy_dn)
    assert _rel(chol_cap, G["gp_chol_cap"]) < 1e-11
    assert float(torch.triu(chol_cap, 1).abs().max()) == 0.0
    assert _rel(gp.log_det_sqrt_gram(dy_du, dy_dn, chol_cap), G["gp_logdet"]) < 1e-11
    assert _rel(gp.log_det_sqrt_gram(dy_du, dy_dn), G["gp_logdet"]) < 1e-11
    assert _rel(gp.lmult_by_jacob_constr(dy_du, dy_dn, vq), G["gp_jv"]) < 1e-12
    assert _rel(gp.rmult_by_jacob_constr(dy_du, dy_dn, wy), G["gp_jtw"]) < 1e-12
    pinv = gp.rmult_by_jacob_constr(dy_du, dy_dn, gp.lmult_by_inv_gram(dy_du, dy_dn, chol_cap, wy))
    assert _rel(pinv, G["gp_pinv_w"]) < 1e-10
    ijp = gp.lmult_by_inv_jacob_product(dy_du, dy_dn, G["gp_dy_du_prev"], G["gp_dy_dn_prev"], wy)
    assert _rel(ijp, G["gp_ijp_w"]) < 1e-10
    nsc = gp.rmult_by_jacob_constr(dy_du, dy_dn, gp.lmult_by_inv_gram(
        dy_du, dy_dn, chol_cap, gp.lmult_by_jacob_constr(dy_du, dy_dn, vq)))
    assert _rel(torch.as_tensor(vq, device="cuda") - nsc, G["gp_proj_v"]) < 1e-10


@pytest.mark.parametrize("dim_y,dim_u", [(1, 1), (5, 2), (40, 16), (129, 3), (260, 7)])
def test_gaussian_process_block_algebra_matches_restatement(dim_y, dim_u):
    import manifold_lifting_b200 as mlb
    from mlift_oracle import gp as ogp

    gp = mlb.gaussian_process
    rng = np.random.default_rng(7 * dim_y + dim_u)
    n = 5
    a, b = rng.standard_normal((2, n, dim_y, dim_y))
    covar = a @ a.transpose(0, 2, 1) / dim_y + np.eye(dim_y)
    L = gp.cholesky(covar)
    assert _rel(L, np.linalg.cholesky(covar)) < 1e-12
    Lr = np.linalg.cholesky(b @ b.transpose(0, 2, 1) / dim_y + np.eye(dim_y))
    Ln = L.cpu().numpy()
    du, dur = rng.standard_normal((2, n, dim_y, dim_u))
    vq, wy = rng.standard_normal((n, dim_u + dim_y)), rng.standard_normal((n, dim_y))
    (cc,) = gp.decompose_gram(du, Ln)
    ref = [ogp.decompose_gram(du[i], Ln[i])[0] for i in range(n)]
    assert _rel(cc, np.stack(ref)) < 1e-11
    assert _rel(gp.lmult_by_inv_gram(du, Ln, cc, wy),
                np.stack([ogp.lmult_by_inv_gram(du[i], Ln[i], ref[i], wy[i]) for i in range(n)])) < 1e-10
    assert _rel(gp.lmult_by_inv_jacob_product(du, Ln, dur, Lr, wy),
                np.stack([ogp.lmult_by_inv_jacob_product(du[i], Ln[i], dur[i], Lr[i], wy[i])
                          for i in range(n)])) < 1e-9
    assert _rel(gp.lmult_by_jacob_constr(du, Ln, vq),
                np.stack([ogp.lmult_by_jacob_constr(du[i], Ln[i], vq[i]) for i in range(n)])) < 1e-12
    assert _rel(gp.rmult_by_jacob_constr(du, Ln, wy),
                np.stack([ogp.rmult_by_jacob_constr(du[i], Ln[i], wy[i]) for i in range(n)])) < 1e-12
    assert _rel(gp.log_det_sqrt_gram(du, Ln),
                np.array([ogp.log_det_sqrt_gram(du[i], Ln[i]) for i in range(n)])) < 1e-11
    # Gram (Gram^-1 w) = w through the Jacobian products: J J^T x
    x = gp.lmult_by_inv_gram(du, Ln, cc, wy)
    back = gp.lmult_by_jacob_constr(du, Ln, gp.rmult_by_jacob_constr(du, Ln, x))
    assert _rel(back, wy) < 1e-9
    # a covariance that is not positive definite poisons its chain only
    bad = covar.copy()
    bad[2] = -np.eye(dim_y)
    Lb = gp.cholesky(bad)
    isn = torch.isnan(Lb).flatten(1).any(1).cpu().numpy()
    assert isn[2] and not isn[[0, 1, 3, 4]].any()
