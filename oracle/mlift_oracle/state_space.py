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
        vct - dc_du @ sla.cho_solve((chol_cap_mtx, True), dc_du.T @ tridiagonal_solve(a, b, a, vct), check_finite=False),
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


# ---- stochastic-volatility model (mlift/example_models/stochastic_volatility.py), the
# `unbounded` prior parametrisation (prior.py:17-42): mu = u0 (normal(0, 1), real support: no
# transform), sigma = exp(u1) (half_normal(1), transforms.py:43-55), phi = -1 + 2 expit(u2)
# (uniform(-1, 1), transforms.py:72-84)


def sv_generate_params(u):
    e = 1.0 / (1.0 + np.exp(-u[2]))
    params = {"μ": u[0], "σ": np.exp(u[1]), "ϕ": -1.0 + 2.0 * e}
    dparams_du = {"μ": 1.0, "σ": np.exp(u[1]), "ϕ": 2.0 * e * (1.0 - e)}
    return params, dparams_du


def sv_constr_split(u, v, n, y):  # stochastic_volatility.py:71-89
    params, _ = sv_generate_params(u)
    x = 2 * np.log(y / n)
    return np.concatenate((
        (params["μ"] + (params["σ"] / (1 - params["ϕ"] ** 2) ** 0.5) * v[0] - x[0])[None],
        params["μ"] + params["ϕ"] * (x[:-1] - params["μ"]) + params["σ"] * v[1:] - x[1:],
    )), x  # fmt: skip


def sv_jacob_constr_split_blocks(u, v, n, y):  # stochastic_volatility.py:92-141
    dim_y = y.shape[0]
    params, dparams_du = sv_generate_params(u)
    x = 2 * np.log(y / n)
    dx_dy = 2 / y
    one_minus_phi_sq = 1 - params["ϕ"] ** 2
    sqrt_one_minus_phi_sq = one_minus_phi_sq**0.5
    v_0_over = v[0] / sqrt_one_minus_phi_sq
    x_minus_mu = x[:-1] - params["μ"]
    dc_du = np.stack((
        np.concatenate((np.array([dparams_du["μ"]]),
                        (1 - params["ϕ"]) * dparams_du["μ"] * np.ones(dim_y - 1))),
        np.concatenate((np.array([dparams_du["σ"] * v_0_over]), dparams_du["σ"] * v[1:])),
        np.concatenate((np.array([dparams_du["ϕ"] * params["ϕ"] * params["σ"] * v_0_over
                                  / one_minus_phi_sq]), x_minus_mu * dparams_du["ϕ"])),
    ), 1)  # fmt: skip
    dc_dv = np.concatenate((np.array([params["σ"] / sqrt_one_minus_phi_sq]),
                            params["σ"] * np.ones(dim_y - 1)))  # fmt: skip
    dc_dn = 2 / n, -2 * params["ϕ"] / n[:-1]
    c = np.concatenate((
        np.array([params["μ"] + params["σ"] * v_0_over - x[0]]),
        params["μ"] + params["ϕ"] * x_minus_mu + params["σ"] * v[1:] - x[1:],
    ))  # fmt: skip
    return (dc_du, dc_dv, dc_dn, dx_dy), c


def sv_split(q, dim_y):
    return q[:3], q[3 : 3 + dim_y], q[3 + dim_y :]


def sv_generate_y(u, v, n):
    """Forward simulation (stochastic_volatility.py:29-38 through
    construct_state_space_model_generators): x_0, x_t, y_t = exp(x_t / 2) n_t."""
    params, _ = sv_generate_params(u)
    x = np.empty(v.shape[0])
    x[0] = params["μ"] + (params["σ"] / (1 - params["ϕ"] ** 2) ** 0.5) * v[0]
    for t in range(1, v.shape[0]):
        x[t] = params["μ"] + params["ϕ"] * (x[t - 1] - params["μ"]) + params["σ"] * v[t]
    return np.exp(x / 2) * n


def maximum_norm(x):  # solvers.py:6-8
    return np.max(np.abs(x))


def _keep_iterating(i, norm_delta_q, error, c_tol, p_tol, div_tol, max_iters):
    # systems.py:215-227
    diverged = error > div_tol or np.isnan(error)
    converged = error < c_tol and norm_delta_q < p_tol
    return not (i >= max_iters or diverged or converged)


def sv_quasi_newton_projection(q, jac_prev, gram_prev, dt, y, c_tol=1e-9, p_tol=1e-8,
                               div_tol=1e10, max_iters=50):  # fmt: skip
    """systems.py:190-228 with the state-space closures; returns (q, mu / dt, i, |dq|, err)."""
    dim_y = y.shape[0]
    q, mu, i, norm_delta_q, error = q.copy(), np.zeros_like(q), 0, np.inf, -1.0
    while _keep_iterating(i, norm_delta_q, error, c_tol, p_tol, div_tol, max_iters):
        c = sv_constr_split(*sv_split(q, dim_y), y)[0]
        error = maximum_norm(c)
        delta_q = rmult_by_jacob_constr(*jac_prev, lmult_by_inv_gram(*jac_prev, *gram_prev, c))
        mu, q, i = mu + delta_q, q - delta_q, i + 1
        norm_delta_q = maximum_norm(delta_q)
    return q, mu / dt, i, norm_delta_q, error


def sv_newton_projection(q, jac_prev, dt, y, c_tol=1e-9, p_tol=1e-8, div_tol=1e10,
                         max_iters=50):  # fmt: skip
    """systems.py:230-270 with the state-space closures."""
    dim_y = y.shape[0]
    q, mu, i, norm_delta_q, error = q.copy(), np.zeros_like(q), 0, np.inf, -1.0
    while _keep_iterating(i, norm_delta_q, error, c_tol, p_tol, div_tol, max_iters):
        jac, c = sv_jacob_constr_split_blocks(*sv_split(q, dim_y), y)
        error = maximum_norm(c)
        delta_q = rmult_by_jacob_constr(
            *jac_prev, lmult_by_inv_jacob_product(*jac, *jac_prev, c))
        mu, q, i = mu + delta_q, q - delta_q, i + 1
        norm_delta_q = maximum_norm(delta_q)
    return q, mu / dt, i, norm_delta_q, error


def sv_observation_noise(rng, dim_y):
    """Synthetic observation noise kept away from zero (|n| >= 0.2): the Jacobian has 2 / n on
    its diagonal, and an observation with n ~ 1e-4 makes the test problem ill-conditioned
    for every implementation alike."""
    z = rng.normal(size=dim_y)
    return np.where(z < 0, -1.0, 1.0) * (0.2 + np.abs(z))


def sv_on_manifold_states(rng, n_chain, dim_y, y=None):
    """Initial states on the manifold: u around typical values, v ~ N(0, I), observations
    generated once from a reference chain, n = y / exp(x / 2) per chain."""
    if y is None:
        u_ref = np.array([-0.5, np.log(0.3), 2.0])
        y = sv_generate_y(u_ref, rng.normal(size=dim_y), sv_observation_noise(rng, dim_y))
    qs = np.empty((n_chain, 3 + 2 * dim_y))
    for i in range(n_chain):
        u = np.array([-0.5, np.log(0.3), 2.0]) + 0.2 * rng.normal(size=3)
        v = rng.normal(size=dim_y)
        params, _ = sv_generate_params(u)
        x = np.empty(dim_y)
        x[0] = params["μ"] + (params["σ"] / (1 - params["ϕ"] ** 2) ** 0.5) * v[0]
        for t in range(1, dim_y):
            x[t] = params["μ"] + params["ϕ"] * (x[t - 1] - params["μ"]) + params["σ"] * v[t]
        qs[i] = np.concatenate((u, v, y / np.exp(x / 2)))
    return qs, y


def sv_constrained_position_step(q, p, dt, y, solver="newton", reverse_check_tol=2e-8,
                                 c_tol=1e-9, p_tol=1e-8, div_tol=1e10, max_iters=50):
    """Position half of Mici 0.2.1's ConstrainedLeapfrogIntegrator step (restated, SURVEY
    Appendix A) with the mlift solver wrappers' contract (solvers.py:56-95) on the
    stochastic-volatility system.  Returns (pos, mom, status, iters_fwd, iters_rev) with
    status 0 ok / 1 diverged / 2 not converged / 3 non-reversible; on a forward failure the
    remaining fields are those computed so far."""
    dim_y = y.shape[0]

    def project(q_, jac_, dt_):
        if solver == "newton":
            return sv_newton_projection(q_, jac_, dt_, y, c_tol, p_tol, div_tol, max_iters)
        return sv_quasi_newton_projection(q_, jac_, decompose_gram(*jac_), dt_, y, c_tol, p_tol,
                                          div_tol, max_iters)  # fmt: skip

    def classify(err, nd):
        if err > div_tol or np.isnan(err):
            return 1
        return 0 if (err < c_tol and nd < p_tol) else 2

    jac0, _ = sv_jacob_constr_split_blocks(*sv_split(q, dim_y), y)
    q1, mu, it_f, nd, err = project(q + dt * p, jac0, dt)
    status = classify(err, nd)
    if status:
        return q1, p - mu, status, it_f, 0
    jac1, _ = sv_jacob_constr_split_blocks(*sv_split(q1, dim_y), y)
    gram1 = decompose_gram(*jac1)
    p1 = p - mu
    p1 = p1 - rmult_by_jacob_constr(
        *jac1, lmult_by_inv_gram(*jac1, *gram1, lmult_by_jacob_constr(*jac1, p1)))
    qb, _, it_r, nd, err = project(q1 - dt * p1, jac1, -dt)
    status = classify(err, nd)
    if status == 0 and not maximum_norm(qb - q) <= reverse_check_tol:
        status = 3
    return q1, p1, status, it_f, it_r


# ---- gradient of log_det_sqrt_gram for the stochastic-volatility model.  The reference
# takes it by reverse-mode autodiff of systems.py:1233-1243 (`jax.value_and_grad`, jit table
# systems.py:332-334); restated here in closed form:
#   d(1/2 log det G) = tr(G^-1 J dJ^T) = sum over the Jacobian's non-zero pattern of
#   (G^-1 J)_{tj} dJ_{tj},   J = [dc_du, diag(dc_dv), bidiag(dc_dn0, dc_dn1)],
# with the derivatives of the entries of stochastic_volatility.py:92-141 taken by hand
# (dx_dy = 2 / y does not depend on q).  Dense G^-1: this is the small-size checker; pinned
# to the reference's autodiff by tests/golden/ref_sv.npz (`closed_grad_logdet`).
def sv_param_derivatives(u):
    e = 1.0 / (1.0 + np.exp(-u[2]))
    sigma, phi = np.exp(u[1]), -1.0 + 2.0 * e
    dphi = 2.0 * e * (1.0 - e)
    return {"mu": u[0], "sigma": sigma, "phi": phi, "dphi": dphi,
            "ddphi": dphi * (1.0 - 2.0 * e)}  # fmt: skip


def sv_grad_log_det_sqrt_gram(q, y):
    dim_y = y.shape[0]
    u, v, n = sv_split(q, dim_y)
    (dc_du, dc_dv, (dn0, dn1), dx_dy), _ = sv_jacob_constr_split_blocks(u, v, n, y)
    jac = dense_jacobian(dc_du, dc_dv, (dn0, dn1), dx_dy)
    ginv = np.linalg.inv(jac @ jac.T)
    gd, go = np.diag(ginv), np.diag(ginv, 1)
    ru = ginv @ dc_du  # weights of the dc_du entries
    wdv = gd * dc_dv  # ... of dc_dv
    w0 = gd * dn0 + np.concatenate((go * dn1, [0.0]))  # ... of dc_dn0
    w1 = go * dn0[:-1] + gd[1:] * dn1  # ... of dc_dn1
    p = sv_param_derivatives(u)
    mu, sigma, phi, dphi, ddphi = p["mu"], p["sigma"], p["phi"], p["dphi"], p["ddphi"]
    s = np.sqrt(1.0 - phi * phi)
    x = 2.0 * np.log(y / n)
    g_u = np.zeros(3)
    g_v = np.zeros(dim_y)
    g_n = np.zeros(dim_y)
    # rows t >= 1
    g_u[0] = -dphi * ru[1:, 2].sum()
    g_u[1] = sigma * (ru[1:, 1] * v[1:]).sum() + sigma * wdv[1:].sum()
    g_u[2] = (-dphi * ru[1:, 0].sum() + ddphi * (ru[1:, 2] * (x[:-1] - mu)).sum()
              - 2.0 * dphi * (w1 / n[:-1]).sum())  # fmt: skip
    g_v[1:] = sigma * ru[1:, 1]
    g_n[:-1] += ru[1:, 2] * dphi * (-2.0 / n[:-1]) + w1 * 2.0 * phi / n[:-1] ** 2
    g_n += w0 * (-2.0 / n**2)
    # row 0
    v0 = v[0]
    g_u[1] += ru[0, 1] * sigma * v0 / s + ru[0, 2] * dphi * phi * sigma * v0 / s**3 + wdv[0] * sigma / s
    g_u[2] += (ru[0, 1] * sigma * v0 * phi * dphi / s**3
               + ru[0, 2] * sigma * v0 * (ddphi * phi / s**3 + dphi**2 / s**3
                                          + 3.0 * phi**2 * dphi**2 / s**5)
               + wdv[0] * sigma * phi * dphi / s**3)  # fmt: skip
    g_v[0] = ru[0, 1] * sigma / s + ru[0, 2] * dphi * phi * sigma / s**3
    return np.concatenate((g_u, g_v, g_n))


def sv_log_det_sqrt_gram(q, y):  # systems.py:1233-1243 at q
    jac, _ = sv_jacob_constr_split_blocks(*sv_split(q, y.shape[0]), y)
    return log_det_sqrt_gram(*jac)


def sv_leapfrog_step(q, p, dt, y, solver="newton", reverse_check_tol=2e-8, **kw):
    """One step of Mici 0.2.1's ConstrainedLeapfrogIntegrator (restated) on
    the stochastic-volatility system with the standard-normal extended prior
    (systems.py:37-42): A(dt/2) = half kick with dh1_dpos (systems.py:368-379) + cotangent
    projection, B(dt) = sv_constrained_position_step, A(dt/2).  Returns (pos, mom, status,
    iters_fwd, iters_rev, h_init, h_final); a failed step leaves (q, p)."""
    dim_y = y.shape[0]

    def project_cotangent(q_, p_):
        jac, _ = sv_jacob_constr_split_blocks(*sv_split(q_, dim_y), y)
        gram = decompose_gram(*jac)
        return p_ - rmult_by_jacob_constr(
            *jac, lmult_by_inv_gram(*jac, *gram, lmult_by_jacob_constr(*jac, p_)))

    def hamiltonian(q_, p_):
        return 0.5 * q_ @ q_ + sv_log_det_sqrt_gram(q_, y) + 0.5 * p_ @ p_

    h_init = hamiltonian(q, p)
    p_half = project_cotangent(q, p - 0.5 * dt * (q + sv_grad_log_det_sqrt_gram(q, y)))
    q1, p1, status, it_f, it_r = sv_constrained_position_step(
        q, p_half, dt, y, solver=solver, reverse_check_tol=reverse_check_tol, **kw)
    if status:
        return q, p, status, it_f, it_r, h_init, h_init
    p1 = project_cotangent(q1, p1 - 0.5 * dt * (q1 + sv_grad_log_det_sqrt_gram(q1, y)))
    return q1, p1, status, it_f, it_r, h_init, hamiltonian(q1, p1)
