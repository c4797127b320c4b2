"""Shared model / state builders for the tests (oracle side)."""
import numpy as np

import c_oracle
import mlift_oracle  # noqa: F401  (sets torch default dtype)
from mlift_oracle.mici import ChainState, DenseConstrainedEuclideanMetricSystem
from mlift_oracle.models import EightSchools, Toy2D
from mlift_oracle.systems import HierarchicalLatentVariableModelSystem

EIGHT_SCHOOLS_Y = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
EIGHT_SCHOOLS_SIGMA = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])


def synthetic_schools(n_school, seed=202101):
    """sigma_j~U(9,18), y_j = 4.4 + 3.6 N(0,1) + sigma_j N(0,1)."""
    rng = np.random.default_rng(seed)
    sigma = rng.uniform(9.0, 18.0, n_school)
    y = 4.4 + 3.6 * rng.standard_normal(n_school) + sigma * rng.standard_normal(n_school)
    return y, sigma


def hier_case(parametrization="unbounded", y=EIGHT_SCHOOLS_Y, sigma=EIGHT_SCHOOLS_SIGMA):
    model = EightSchools(y, sigma, parametrization)
    system = HierarchicalLatentVariableModelSystem(
        model.generate_y, model.data, model.dim_u, model.extended_prior_neg_log_dens
    )
    cm = c_oracle.OracleModel.hier(y, sigma, parametrization)
    return model, system, cm


def hier_states(model, n, seed=7, scale=1.0):
    """On-manifold positions with moderate u (prior draws can be extreme)."""
    rng = np.random.default_rng(seed)
    Y = model.dim_y
    u = np.stack([rng.normal(0.0, 1.0, n) * scale, rng.normal(0.5, 0.7, n) * scale], 1)
    v = rng.standard_normal((n, Y))
    if model.data["parametrization"] == "normal":
        mu, tau = 5.0 * u[:, 0], None
        import torch

        tau = model.generate_params(torch.as_tensor(u.T), model.data)["tau"].numpy()
    else:
        mu, tau = u[:, 0], np.exp(u[:, 1])
    x = mu[:, None] + tau[:, None] * v
    nn = (model.data["y_obs"].numpy()[None] - x) / model.data["sigma"].numpy()[None]
    return np.concatenate([u, v, nn], 1)


def toy_case(sigma=0.1, y=1.0, coeff=2.0):
    model = Toy2D(sigma, y, coeff)
    system = DenseConstrainedEuclideanMetricSystem(Toy2D.neg_log_dens, model.constr)
    cm = c_oracle.OracleModel.toy(sigma, y, coeff)
    return model, system, cm


def toy_states(model, n, seed=3):
    rng = np.random.default_rng(seed)
    return model.lift(rng.standard_normal((n, 2)))


def tangent_momentum(cm, q, seed=11):
    rng = np.random.default_rng(seed)
    p = rng.standard_normal(q.shape)
    jac, _ = cm.jacob_constr_blocks(q)
    gram = cm.decompose_gram(jac)
    return p - cm.normal_space_component(jac, gram, p)


def fn_data(golden_dir=None):
    """Reference FN data arrays (committed fixture) -> (fixture, y_obs)."""
    import os

    golden_dir = golden_dir or os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(golden_dir, "fitzhugh_nagumo.npz"))
    return g, g["y_mean_obs"] + float(g["obs_noise_std"]) * g["n_obs"]


def hh_data(golden_dir=None):
    """Reference HH constants / observations (committed fixture) -> (fixture, y_obs)."""
    import os

    golden_dir = golden_dir or os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(golden_dir, "hodgkin_huxley.npz"))
    return g, g["x_obs"] + float(g["obs_noise_std"]) * g["n_obs"]


def hh_states(cm, g, n, seed=5, spread=0.02, off_manifold=0.0):
    """On-manifold HH positions ("normal" parametrisation): u = u_obs + spread * N(0, I),
    n = (y - x(u)) / sigma(u), lifted with the CPU oracle."""
    rng = np.random.default_rng(seed)
    u = g["u_obs"][None] + spread * rng.standard_normal((n, 7))
    q = np.concatenate([u, np.zeros((n, cm.Y))], 1)
    q[:, 7:] = -cm.constr(q) / np.exp(u[:, 6:7])
    if off_manifold:
        q[:, 7:] += off_manifold * rng.standard_normal((n, cm.Y))
    return q


FN_LOC = np.array([0.0, 0.0, 0.0, -1.0, -3.0, -2.0, -1.0])
FN_U_TRUE = np.concatenate([np.log([3, 1, 3, 1 / 3, 1 / 15, 1 / 15, 0.1]), [-1.0, 1.0]])


def fn_states(cm, n, seed=5, par="unbounded", spread=0.05, off_manifold=0.0):
    """On-manifold FN positions: u = truth + spread * N(0, I), n = (y - x(u)) / sigma(u)
    (SURVEY 8d config 4), lifted with the CPU oracle."""
    rng = np.random.default_rng(seed)
    u = FN_U_TRUE[None] + spread * rng.standard_normal((n, 9))
    if par == "normal":
        u[:, :7] -= FN_LOC[None]
    q = np.concatenate([u, np.zeros((n, cm.Y))], 1)
    sigma = np.exp(u[:, 6] + (FN_LOC[6] if par == "normal" else 0.0))
    q[:, 9:] = -cm.constr(q) / sigma[:, None]
    if off_manifold:
        q[:, 9:] += off_manifold * rng.standard_normal((n, cm.Y))
    return q


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (1e-300 + np.maximum(np.abs(b), 1.0))))


__all__ = [n for n in dir() if not n.startswith("__")]


def toy_posterior_moments(sigma=0.1, y=1.0, coeff=2.0, half_width=3.0, n_grid=6001):
    """E[theta0^2], E[theta1^2], E[|theta0|], E[|theta1|] under p(theta | y) ∝ N(theta; 0, I)
    N(y; F(theta), sigma^2) by tensor-product quadrature on a fine grid.  The theta-marginal
    of the lifted constrained target IS this posterior (notebook cells around :1580), so
    this is an anchor that depends on neither implementation under test."""
    g = np.linspace(-half_width, half_width, n_grid)
    t0, t1 = g[:, None], g[None, :]
    F = t1**2 + coeff * t0**2 * (t0**2 - 0.5)
    logp = -0.5 * (t0**2 + t1**2) - 0.5 * ((y - F) / sigma) ** 2
    w = np.exp(logp - logp.max())
    z = w.sum()
    return {"t0_sq": float((w * t0**2).sum() / z), "t1_sq": float((w * t1**2).sum() / z),
            "t0_abs": float((w * np.abs(t0)).sum() / z),
            "t1_abs": float((w * np.abs(t1)).sum() / z)}  # fmt: skip


def toy_manifold_inits(n, rng, y=1.0, coeff=2.0):
    """Points on the limiting manifold F(theta) = y with eta = 0 (what the reference
    starts its toy chains from, toy_2d_model.py:119-157; closed form for coeff = 2, y = 1:
    theta1 = +-sqrt(y - coeff theta0^2 (theta0^2 - 1/2)), |theta0| <= 1)."""
    t0 = rng.uniform(-1.0, 1.0, n)
    t1 = np.sqrt(np.maximum(y - coeff * t0**2 * (t0**2 - 0.5), 0.0)) * rng.choice([-1.0, 1.0], n)
    return np.stack([t0, t1, np.zeros(n)], 1)


def mc_standard_error(x):
    """Monte-Carlo standard error of the mean of x[n_chain, n_draw] from the spread of
    the per-chain means (chains are independent)."""
    m = np.asarray(x).mean(1)
    return float(m.std(ddof=1) / np.sqrt(m.size))


__all__ = [n for n in dir() if not n.startswith("__")]


def woodbury_projection_truth(dy_du, dy_dn, v, dps=50):
    """Extended-precision (mpmath, `dps` digits) normal-space component J^T (J J^T)^-1 J v of
    ONE chain for an additive-noise system, J = [dy_du, diag(dy_dn)] (systems.py:564-654),
    with the fp64 Jacobian entries taken as exact data.  Returns (truth as float64 [Q],
    condition number of the capacitance matrix I + A^T D^-1 A)."""
    import mpmath as mp

    mp.mp.dps = dps
    Y, U = dy_du.shape
    A = [[mp.mpf(float(dy_du[i, a])) for a in range(U)] for i in range(Y)]
    dn = [mp.mpf(float(x)) for x in dy_dn]
    vu = [mp.mpf(float(x)) for x in v[:U]]
    vn = [mp.mpf(float(x)) for x in v[U:]]
    rd = [1 / (d * d) for d in dn]  # D^-1
    w = [sum(A[i][a] * vu[a] for a in range(U)) + dn[i] * vn[i] for i in range(Y)]  # J v
    C = mp.matrix(U, U)
    for a in range(U):
        for b in range(a, U):
            s = sum(A[i][a] * rd[i] * A[i][b] for i in range(Y))
            C[a, b] = C[b, a] = s
        C[a, a] += 1
    rhs = mp.matrix([sum(A[i][a] * rd[i] * w[i] for i in range(Y)) for a in range(U)])
    z = mp.lu_solve(C, rhs)
    x = [rd[i] * (w[i] - sum(A[i][a] * z[a] for a in range(U))) for i in range(Y)]  # G^-1 J v
    out = [sum(A[i][a] * x[i] for i in range(Y)) for a in range(U)] + [dn[i] * x[i] for i in range(Y)]
    Cf = np.array([[float(C[a, b]) for b in range(U)] for a in range(U)])
    return np.array([float(t) for t in out]), float(np.linalg.cond(Cf))
