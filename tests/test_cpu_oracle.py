"""CPU tests of the oracle itself (no GPU): the closed-form C++ oracle against the golden
fixtures, against the torch autodiff restatement, and against the invariants of SURVEY 8c.

"Parity unpinned": the reference has no tests for the constrained path and cannot run
here; the only reference-produced numbers are the notebook's Euclidean-leapfrog output.
"""
import os

import numpy as np
import pytest
import torch

import c_oracle as co
from _cases import (ChainState, hier_case, hier_states, rel_err, tangent_momentum, toy_case,
                    toy_states)  # fmt: skip

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("hier", "unbounded"), ("hier", "normal"), ("toy", None)]


def make(kind, par, n, seed=7):
    if kind == "hier":
        omodel, osys, cm = hier_case(par)
        return omodel, osys, cm, hier_states(omodel, n, seed)
    omodel, osys, cm = toy_case(0.1)
    return omodel, osys, cm, toy_states(omodel, n, seed)


def test_notebook_euclidean_leapfrog_known_answer():
    """The reference tree's only numeric known-answer vector (notebook raw lines 692-753):
    pins the toy forward function (factor 2), the posterior gradient and leapfrog order."""
    g = np.load(os.path.join(GOLD, "notebook_euclidean_leapfrog.npz"))
    sigma, y, eps, k = (float(g[x]) for x in ("sigma", "y", "step_size", "coeff"))
    from mlift_oracle.models import Toy2D

    model = Toy2D(sigma, y, k)

    def grad_u(theta):
        t = torch.tensor(theta, requires_grad=True)
        u = (t**2).sum() / 2 + (y - model.forward_func(t)) ** 2 / (2 * sigma**2)
        return torch.autograd.grad(u, t)[0].numpy()

    pos, mom = g["pos"][0].copy(), g["mom"][0].copy()
    for s in (1, 2):
        mom = mom - 0.5 * eps * grad_u(pos)
        pos = pos + eps * mom
        mom = mom - 0.5 * eps * grad_u(pos)
        assert np.allclose(pos, g["pos"][s], rtol=0, atol=2e-7)
        assert np.allclose(mom, g["mom"][s], rtol=0, atol=2e-6)
    # the closed-form C++ toy Jacobian is the same dF/dtheta
    cm = co.OracleModel.toy(sigma, y, k)
    q = model.lift(g["pos"])
    (dy_du, _, dy_dn), c = cm.jacob_constr_blocks(q)
    t = torch.tensor(g["pos"][1], requires_grad=True)
    df = torch.autograd.grad(model.forward_func(t), t)[0].numpy()
    assert np.allclose(dy_du[1, 0], df, rtol=1e-14) and np.allclose(dy_dn, sigma)
    assert np.abs(c).max() < 1e-12  # lifted points are on the manifold


@pytest.mark.parametrize("name,kind,par", [("hier_unbounded", "hier", "unbounded"),
                                           ("hier_normal", "hier", "normal"),
                                           ("toy", "toy", None)])  # fmt: skip
def test_c_oracle_matches_golden_autodiff_steps(name, kind, par):
    """Closed-form C++ oracle vs committed outputs of the autodiff restatement: projected
    positions, momenta, Hamiltonian and per-step (forward, reverse) iteration counts."""
    g = np.load(os.path.join(GOLD, f"autodiff_steps_{name}.npz"))
    _, _, cm, _ = make(kind, par, 1)
    for sid in g["solvers"]:
        r = cm.leapfrog_steps(g["q0"], g["p0"], float(g["dt"]), int(g["n_step"]),
                              co.solver_opts(int(sid)))  # fmt: skip
        assert (r["status"] == 0).all()
        assert rel_err(r["q"], g[f"q_{sid}"]) < 1e-10
        assert rel_err(r["p"], g[f"p_{sid}"]) < 1e-10
        assert rel_err(r["h_final"], g[f"h_{sid}"]) < 1e-10
        its = g[f"iters_{sid}"]  # [chain, step, (fwd, rev)]
        assert (r["iters_fwd"] == its[:, :, 0].sum(1)).all()
        assert (r["iters_rev"] == its[:, :, 1].sum(1)).all()


@pytest.mark.parametrize("kind,par", CASES)
def test_c_oracle_ops_match_autodiff(kind, par):
    """Jacobian blocks, Gram factor, log-det and its gradient (reverse-over-forward in the
    reference, systems.py:332-334) from autodiff vs the closed forms."""
    omodel, osys, cm, q = make(kind, par, 5)
    (dy_du, dy_dv, dy_dn), c = cm.jacob_constr_blocks(q)
    g_ld, v_ld = cm.grad_log_det_sqrt_gram(q)
    v_nld, g_nld = cm.neg_log_dens(q)
    for i in range(q.shape[0]):
        st = ChainState(q[i])
        assert rel_err(c[i], osys.constr(st)) < 1e-12
        if kind == "hier":
            jb = osys.jacob_constr_blocks(st)
            assert rel_err(dy_du[i], jb[0]) < 1e-12
            assert rel_err(dy_dv[i], jb[1]) < 1e-12 and rel_err(dy_dn[i], jb[2]) < 1e-12
        else:
            jac = np.asarray(osys.jacob_constr(st))
            assert rel_err(np.concatenate([dy_du[i].ravel(), dy_dn[i]]), jac.ravel()) < 1e-12
        assert rel_err(v_ld[i], osys.log_det_sqrt_gram(st)) < 1e-12
        assert rel_err(g_ld[i], osys.grad_log_det_sqrt_gram(st)) < 1e-10
        assert rel_err(v_nld[i], osys.neg_log_dens(st)) < 1e-12
        assert rel_err(g_nld[i], osys.grad_neg_log_dens(st)) < 1e-12


@pytest.mark.parametrize("kind,par", CASES)
def test_grad_log_det_vs_finite_differences(kind, par):
    _, _, cm, q = make(kind, par, 4)
    g, _ = cm.grad_log_det_sqrt_gram(q)
    eps = 1e-6
    for k in range(q.shape[1]):
        dq = np.zeros_like(q)
        dq[:, k] = eps
        fd = (cm.log_det_sqrt_gram(q + dq) - cm.log_det_sqrt_gram(q - dq)) / (2 * eps)
        assert np.allclose(fd, g[:, k], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("kind,par", CASES)
def test_projection_invariants(kind, par):
    """on-manifold after projection, J p = 0 after cotangent
    projection, the three mlift solvers and the two Mici solvers land on one point."""
    _, _, cm, q = make(kind, par, 200)
    p = tangent_momentum(cm, q)
    jac, _ = cm.jacob_constr_blocks(q)
    assert np.abs(cm.lmult_by_jacob_constr(jac, p)).max() < 1e-10
    dt = 0.1
    outs = []
    for s in range(5):
        r = cm.solve_projection(q + dt * p, q, dt, co.solver_opts(s))
        ok = r["status"] == 0
        assert ok.mean() > 0.9
        assert np.abs(cm.constr(r["q"]))[ok].max() < 1e-8
        # the correction lies in span(J_prev^T): q_out = q_try - dt * mu
        assert rel_err((q + dt * p - dt * r["mu"])[ok], r["q"][ok]) < 1e-12
        outs.append((ok, r["q"]))
    ok = np.all([o[0] for o in outs], 0)
    for _, qo in outs[1:]:
        assert np.abs(qo - outs[0][1])[ok].max() < 1e-6


@pytest.mark.parametrize("kind,par", CASES)
def test_step_reversibility_and_energy(kind, par):
    """forward then momentum-flipped steps return to the start;
    energy error is O(dt^2)."""
    _, _, cm, q = make(kind, par, 200)
    p = tangent_momentum(cm, q)
    opts = co.solver_opts(co.SOLVER_NEWTON)
    errs = []
    for dt in (0.1, 0.05):
        f = cm.leapfrog_steps(q, p, dt, 4, opts)
        ok = f["status"] == 0
        b = cm.leapfrog_steps(f["q"][ok], -f["p"][ok], dt, 4, opts)
        ok2 = b["status"] == 0
        assert np.abs(b["q"][ok2] - q[ok][ok2]).max() < 1e-6
        assert np.abs(b["p"][ok2] + p[ok][ok2]).max() < 1e-6
        errs.append(np.median(np.abs(f["h_final"] - f["h_init"])[ok]))
    assert errs[1] < errs[0] / 2.5


def test_woodbury_equals_dense_gram():
    """the Woodbury factorisation (systems.py:749-794) agrees with the
    dense Gram matrix assembled from the same Jacobian blocks."""
    _, _, cm, q = make("hier", "unbounded", 20)
    (dy_du, dy_dv, dy_dn), _ = cm.jacob_constr_blocks(q)
    chol, d = cm.decompose_gram((dy_du, dy_dv, dy_dn))
    rng = np.random.default_rng(0)
    w = rng.standard_normal((20, cm.Y))
    x = cm.lmult_by_pinv_jacob_constr((dy_du, dy_dv, dy_dn), (chol, d), w)
    for i in range(20):
        jfull = np.concatenate([dy_du[i], np.diag(dy_dv[i]), np.diag(dy_dn[i])], 1)
        gram = jfull @ jfull.T
        assert rel_err(x[i], jfull.T @ np.linalg.solve(gram, w[i])) < 1e-11
        sign, logdet = np.linalg.slogdet(gram)
        assert abs(0.5 * logdet - cm.log_det_sqrt_gram(q[i : i + 1])[0]) < 1e-11


def test_philox_stream_reference_values():
    """Counter-based RNG: uniforms in (0,1), normals with unit variance, streams differ
    by chain and iteration (the sharding-independence contract of SURVEY 8e)."""
    z = np.array([co.philox_normals(1, c, 0, 18) for c in range(4000)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    u = np.array([co.philox_uniform(1, c, 0) for c in range(4000)])
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    assert not np.allclose(co.philox_normals(1, 5, 0, 4), co.philox_normals(1, 5, 1, 4))
    assert not np.allclose(co.philox_normals(1, 5, 0, 4), co.philox_normals(1, 6, 0, 4))


def test_oracle_sampler_reproduces_toy_posterior_moments():
    """Statistical anchor for the oracle (the reference has no bit-wise golden vector for
    the constrained path): static constrained HMC of the C++ oracle on the toy
    model against tensor-product quadrature of p(theta | y), within Monte-Carlo error."""
    from _cases import mc_standard_error, toy_manifold_inits, toy_posterior_moments

    sigma = 0.1
    cm = co.OracleModel.toy(sigma, 1.0, 2.0)
    q = toy_manifold_inits(512, np.random.default_rng(2))
    opts = co.solver_opts(co.SOLVER_NEWTON)
    n_warm, n_main = 300, 500
    tr = np.empty((q.shape[0], n_main, 2))
    acc = []
    for it in range(n_warm + n_main):
        r = cm.hmc_transition(q, 0.1, 20, opts, seed=4, iteration=it)
        q = r["q"]
        if it >= n_warm:
            tr[:, it - n_warm] = q[:, :2]
            acc.append(r["accept_stat"].mean())
    assert np.mean(acc) > 0.6
    want = toy_posterior_moments(sigma)
    got = {"t0_sq": tr[..., 0] ** 2, "t1_sq": tr[..., 1] ** 2,
           "t0_abs": np.abs(tr[..., 0]), "t1_abs": np.abs(tr[..., 1])}  # fmt: skip
    for k, x in got.items():
        se = mc_standard_error(x)
        assert se < 2e-3 and abs(x.mean() - want[k]) < 5 * se, (k, x.mean(), want[k], se)


def test_oracle_dynamic_sampler_reproduces_toy_posterior_and_notebook_accept_rate():
    """The recursive dynamic multinomial sampler of the oracle (restating Mici's
    DynamicMultinomialHMC) at the notebook's step size 0.1: accept statistic 0.995-0.996
    (notebook raw lines 2565-2684 report 0.995 / 0.996 for the constrained chains),
    no integrator errors, posterior moments against quadrature."""
    from _cases import mc_standard_error, toy_manifold_inits, toy_posterior_moments

    sigma = 0.1
    cm = co.OracleModel.toy(sigma, 1.0, 2.0)
    q = toy_manifold_inits(384, np.random.default_rng(2))
    opts = co.solver_opts(co.SOLVER_NEWTON)
    n_warm, n_main = 100, 300
    tr = np.empty((q.shape[0], n_main, 2))
    acc, bad = [], []
    for it in range(n_warm + n_main):
        r = cm.nuts_transition(q, 0.1, opts, seed=4, iteration=it)
        q = r["q"]
        if it >= n_warm:
            tr[:, it - n_warm] = q[:, :2]
            acc.append(r["accept_stat"].mean()), bad.append((r["status"] != 0).mean())
    assert 0.990 < np.mean(acc) < 0.999 and np.mean(bad) < 2e-3
    want = toy_posterior_moments(sigma)
    for j, k in enumerate(("t0_sq", "t1_sq")):
        x = tr[..., j] ** 2
        se = mc_standard_error(x)
        assert se < 3e-3 and abs(x.mean() - want[k]) < 5 * se, (k, x.mean(), want[k], se)


def test_oracle_dynamic_sampler_extra_subtree_checks_and_one_sided_divergence():
    """Mici's defaults for DynamicMultinomialHMC (utils.py:286-288 passes none of them):
    `do_extra_subtree_checks=True` only ADDS termination tests, so from the same state with
    the same draws a tree built with them is never larger than without, is identical when
    no extra test fires, and on the funnel-shaped eight-schools posterior some trees do
    stop earlier; the one-sided divergence test accepts any energy decrease."""
    y = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
    s = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])
    cm = co.OracleModel.hier(y, s)
    rng = np.random.default_rng(11)
    n = 1024
    u = np.stack([rng.normal(4.0, 3.0, n), rng.normal(1.0, 0.7, n)], 1)
    v = rng.standard_normal((n, 8))
    q = np.concatenate([u, v, (y[None] - u[:, :1] - np.exp(u[:, 1:2]) * v) / s[None]], 1)
    on = co.solver_opts(co.SOLVER_NEWTON)
    off = co.solver_opts(co.SOLVER_NEWTON, flags=co.NUTS_NO_EXTRA_SUBTREE_CHECKS)
    fewer = 0
    for it in range(6):
        a = cm.nuts_transition(q, 0.5, on, max_depth=8, seed=3, iteration=it)
        b = cm.nuts_transition(q, 0.5, off, max_depth=8, seed=3, iteration=it)
        assert (a["n_step"] <= b["n_step"]).all()
        # (equal steps with a smaller depth: the last subtree's final merge was refused)
        same = (a["n_step"] == b["n_step"]) & (a["depth"] == b["depth"])
        assert np.array_equal(a["q"][same], b["q"][same])
        assert np.array_equal(a["accept_stat"][same], b["accept_stat"][same])
        fewer += int((~same).sum())
        q = a["q"]
    assert 0 < fewer < 0.2 * 6 * n, fewer
    # a (huge) energy DECREASE is divergent only under the two-sided test
    two = co.solver_opts(co.SOLVER_NEWTON, flags=co.NUTS_TWO_SIDED_DELTA_H)
    a = cm.nuts_transition(q, 0.5, on, max_depth=6, seed=5, max_delta_h=0.05)
    b = cm.nuts_transition(q, 0.5, two, max_depth=6, seed=5, max_delta_h=0.05)
    H_DIV = 4
    assert (b["status"] == H_DIV).sum() > (a["status"] == H_DIV).sum() > 0


def test_tridiagonal_restatement_matches_dense_linear_algebra():
    """oracle/mlift_oracle/linalg.py (the scans of mlift/linalg.py:46-114 as loops) against
    dense NumPy: np.linalg.solve and slogdet of the assembled matrix, including dim 1 and 2
    and a non-symmetric system (a != c, as in lmult_by_inv_jacob_product, systems.py:1212-1231)."""
    from mlift_oracle import linalg as ol

    rng = np.random.default_rng(0)
    for dim in (1, 2, 3, 10, 200):
        a, c = rng.normal(size=dim - 1), rng.normal(size=dim - 1)
        b = 2.5 + np.abs(rng.normal(size=dim)) + np.pad(np.abs(a), (1, 0)) + np.pad(np.abs(c), (0, 1))
        d = rng.normal(size=dim)
        T = np.diag(b) + np.diag(a, -1) + np.diag(c, 1)
        x = ol.tridiagonal_solve(a, b, c, d)
        assert np.abs(x - np.linalg.solve(T, d)).max() < 1e-12 * max(1.0, np.abs(x).max())
        S = np.diag(b) + np.diag(a, -1) + np.diag(a, 1)
        sign, ld = np.linalg.slogdet(S)
        assert sign > 0 and abs(ol.tridiagonal_pos_def_log_det(a, b) - ld) < 1e-10 * max(1.0, abs(ld))


def test_state_space_restatement_matches_dense_linear_algebra():
    """oracle/mlift_oracle/state_space.py (the closures of
    PartiallyInvertibleStateSpaceModelSystem, mlift/systems.py:1161-1243) against the dense
    Jacobian the blocks stand for: products, (J J^T)^{-1}, (J_l J_r^T)^{-1} and
    log det(J J^T) / 2 - sum log|dx_dy|, including dim_y 1 and 2."""
    from mlift_oracle import state_space as oss

    rng = np.random.default_rng(1)
    for dim_y, dim_u in ((1, 1), (2, 2), (7, 3), (60, 5), (300, 3)):
        J, Jr = oss.random_blocks(rng, dim_y, dim_u), oss.random_blocks(rng, dim_y, dim_u)
        Jd, Jrd = oss.dense_jacobian(*J), oss.dense_jacobian(*Jr)
        vq, vy = rng.normal(size=dim_u + 2 * dim_y), rng.normal(size=dim_y)
        assert np.abs(oss.lmult_by_jacob_constr(*J, vq) - Jd @ vq).max() < 1e-13
        assert np.abs(oss.rmult_by_jacob_constr(*J, vy) - vy @ Jd).max() < 1e-13
        a, b, chol = oss.decompose_gram(*J)
        G = Jd @ Jd.T
        T = np.diag(b) + (np.diag(a, -1) + np.diag(a, 1) if dim_y > 1 else 0.0)
        assert np.abs(T + J[0] @ J[0].T - G).max() < 1e-12 * np.abs(G).max()
        x = oss.lmult_by_inv_gram(*J, a, b, chol, vy)
        assert np.abs(x - np.linalg.solve(G, vy)).max() < 1e-11 * max(1.0, np.abs(x).max())
        x = oss.lmult_by_inv_jacob_product(*J, *Jr, vy)
        ref = np.linalg.solve(Jd @ Jrd.T, vy)
        assert np.abs(x - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        want = 0.5 * np.linalg.slogdet(G)[1] - np.log(np.abs(J[3])).sum()
        assert abs(oss.log_det_sqrt_gram(*J) - want) < 1e-10 * max(1.0, abs(want))


def test_stochastic_volatility_restatement_jacobian_and_projection():
    """oracle/mlift_oracle/state_space.py, stochastic-volatility part
    (mlift/example_models/stochastic_volatility.py:71-141): the blocks are the Jacobian of
    constr_split (central differences), simulated states sit on the manifold, and both
    projection loops (systems.py:190-270) return to it from a cotangent step."""
    from mlift_oracle import state_space as oss

    rng = np.random.default_rng(2)
    dim_y = 25
    qs, y = oss.sv_on_manifold_states(rng, 3, dim_y)
    for q in qs:
        c, _ = oss.sv_constr_split(*oss.sv_split(q, dim_y), y)
        assert np.abs(c).max() < 1e-12
        jac, c2 = oss.sv_jacob_constr_split_blocks(*oss.sv_split(q, dim_y), y)
        assert np.abs(c - c2).max() < 1e-14
        Jd = oss.dense_jacobian(*jac)
        fd = np.empty_like(Jd)
        for j in range(q.size):
            e = np.zeros_like(q)
            e[j] = 1e-6 * max(1.0, abs(q[j]))
            fd[:, j] = (oss.sv_constr_split(*oss.sv_split(q + e, dim_y), y)[0]
                        - oss.sv_constr_split(*oss.sv_split(q - e, dim_y), y)[0]) / (2 * e[j])  # fmt: skip
        assert np.abs(Jd - fd).max() < 1e-6 * np.abs(Jd).max()
        gram = oss.decompose_gram(*jac)
        p = rng.normal(size=q.size)
        p -= oss.rmult_by_jacob_constr(
            *jac, oss.lmult_by_inv_gram(*jac, *gram, oss.lmult_by_jacob_constr(*jac, p)))
        q1 = q + 0.05 * p
        qa, _, ia, _, ea = oss.sv_quasi_newton_projection(q1, jac, gram, 0.05, y)
        qb, _, ib, _, eb = oss.sv_newton_projection(q1, jac, 0.05, y)
        assert ia < 50 and ib < 50 and ib <= ia and max(ea, eb) < 1e-9
        for qq in (qa, qb):
            assert np.abs(oss.sv_constr_split(*oss.sv_split(qq, dim_y), y)[0]).max() < 1e-9
        # same retraction direction (span of J_prev^T): the two solvers land on the same point
        assert np.abs(qa - qb).max() < 1e-7
