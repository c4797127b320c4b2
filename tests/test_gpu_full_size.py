"""Per-chain oracle parity AT THE BASELINE SIZES (not just invariants): 65,536 eight-schools
chains (both parametrisations, the three mlift solvers), 65,536 toy chains (Mici's own
solvers), 16,384 FitzHugh-Nagumo chains and 1,024 Hodgkin-Huxley chains - CUDA path vs the
C++ oracle (itself pinned to the reference's source, tests/test_cpu_ref_pinning.py) chain by
chain: status, completed steps, iteration counts, positions, momenta, Hamiltonians.

Iteration counts must be EQUAL except on "edge" chains whose deciding norm sits within
rounding of a tolerance (1-ulp differences between CPU and GPU arithmetic move the exit by
one iteration); the permitted edge fraction is stated per test, and edge chains must still
end at the same point to solver tolerance.
"""
import numpy as np
import pytest
import torch

import c_oracle as co
from _cases import (fn_data, fn_states, hh_data, hh_states, rel_err, synthetic_schools,
                    tangent_momentum)  # fmt: skip

pytestmark = pytest.mark.gpu
TOL = 1e-10


def cpu(t):
    return t.detach().cpu().numpy()


def scaled_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    return float((np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1.0)).max())


def compare(out, ref, tol_q, tol_p, tol_h, max_edge, err=rel_err, edge_tol=1e-7, explained=None):
    """Chain-by-chain comparison; returns the fraction of edge chains.  `explained`: chains
    whose iteration path is known not to be reproducible (line search decided by rounding);
    `max_edge` then bounds the edge chains OUTSIDE that set."""
    st, n_done = cpu(out.status), cpu(out.n_done)
    itf, itr = cpu(out.iters_fwd), cpu(out.iters_rev)
    same_it = (itf == ref["iters_fwd"]) & (itr == ref["iters_rev"])
    same_st = (st == ref["status"]) & (n_done == ref["n_done"])
    edge = ~(same_it & same_st)
    unexplained = edge if explained is None else edge & ~explained
    assert unexplained.mean() <= max_edge, (
        f"{unexplained.sum()} of {edge.size} chains differ in counts / status"
        + ("" if explained is None else f" outside the {explained.sum()} flagged chains"))
    # an edge chain differs by at most one iteration per solve and, when both completed,
    # sits at the same point to solver tolerance
    both = edge & (st == 0) & (ref["status"] == 0)
    if both.any():
        assert np.abs(cpu(out.pos)[both] - ref["q"][both]).max() < edge_tol
    ok = ~edge & (ref["status"] == 0)
    assert ok.mean() > 0.9
    assert err(cpu(out.pos)[ok], ref["q"][ok]) < tol_q
    assert err(cpu(out.mom)[ok], ref["p"][ok]) < tol_p
    assert rel_err(cpu(out.h_init), ref["h_init"]) < tol_h
    assert rel_err(cpu(out.h_final)[ok], ref["h_final"][ok]) < tol_h
    bad = ~edge & (ref["status"] != 0)  # failed chains stay at their last completed step
    if bad.any():
        assert err(cpu(out.pos)[bad], ref["q"][bad]) < tol_q
    return float(edge.mean())


def integrator(mlb, gs, dt, solver):
    names = {v: k for k, v in mlb.solvers.SOLVER_IDS.items()}
    return mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, dt, projection_solver=names[solver],
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))


@pytest.mark.parametrize("par", ["unbounded", "normal"])
@pytest.mark.parametrize("solver", [co.SOLVER_QUASI_NEWTON, co.SOLVER_NEWTON, co.SOLVER_NEWTON_LS])
def test_eight_schools_65536_chains_match_oracle_per_chain(par, solver):
    """BASELINE config 3 at full size, the bench's inputs (prior-like latents incl. the
    half-Cauchy tail, clipped as bench.py does), 3 steps of 0.1."""
    import manifold_lifting_b200 as mlb

    y, sigma = synthetic_schools(8)
    n = 65536
    rng = np.random.default_rng([202101, 7])
    u = np.stack([rng.normal(0.0, 5.0 if par == "unbounded" else 1.0, n),
                  np.clip(np.log(np.abs(rng.standard_cauchy(n) * 5.0)), -2.0, 3.0)
                  if par == "unbounded" else rng.normal(0.0, 1.0, n)], 1)  # fmt: skip
    v = rng.standard_normal((n, 8))
    gm = mlb.models.EightSchoolsModel(y, sigma, par)
    gs = mlb.HierarchicalLatentVariableModelSystem(gm, gm.data, gm.dim_u)
    pr = gm.generate_params(u)
    q = np.concatenate([u, v, (y[None] - pr["μ"][:, None] - pr["τ"][:, None] * v) / sigma[None]], 1)
    cm = co.OracleModel.hier(y, sigma, par)
    co.use_all_cores()
    p = tangent_momentum(cm, q)
    dt, n_step = 0.1, 3
    ref = cm.leapfrog_steps(q, p, dt, n_step, co.solver_opts(solver))
    # line search: a backtracking comparison |c(q + a d)| > |c(q)| decided by rounding (both
    # norms residual noise, or equal to 1e-6 relative: flagged by the oracle) changes the
    # halving count and can change the iteration path; measured, that moves the path of
    # < 0.1 % of the chains, so ALL solvers are held to the same 0.5 % bound on edge chains
    amb = cm.last_line_search_ambiguous(n) if solver == co.SOLVER_NEWTON_LS else None
    out = integrator(mlb, gs, dt, solver).steps(mlb.states.ChainState(q, p), n_step)
    edge = compare(out, ref, TOL, TOL, TOL, 0.005)
    print(f"eight-schools {par} solver {solver}: edge chains {edge:.5f}, "
          f"failed (both) {(ref['status'] != 0).mean():.4f}"
          + ("" if amb is None else f", a line-search comparison decided by rounding in "
             f"{amb.mean():.4f} of the chains"))


@pytest.mark.parametrize("solver", [co.SOLVER_MICI_NEWTON, co.SOLVER_MICI_QUASI_NEWTON])
def test_toy_65536_chains_match_oracle_per_chain(solver):
    """BASELINE config 2 at full size: theta ~ N(0, I) lifted, sigma = 0.1, Mici's own
    solvers (toy_2d_model.py:170-174), 3 steps of 0.1."""
    import manifold_lifting_b200 as mlb

    n = 65536
    rng = np.random.default_rng([202101, 3])
    gm = mlb.models.Toy2DModel(0.1, 1.0, 2.0)
    gs = mlb.IndependentAdditiveNoiseModelSystem(gm, gm.data, gm.dim_u)
    q = gm.lift(rng.standard_normal((n, 2)))
    cm = co.OracleModel.toy(0.1, 1.0, 2.0)
    co.use_all_cores()
    p = tangent_momentum(cm, q)
    dt, n_step = 0.1, 3
    ref = cm.leapfrog_steps(q, p, dt, n_step, co.solver_opts(solver))
    out = integrator(mlb, gs, dt, solver).steps(mlb.states.ChainState(q, p), n_step)
    # strongly curved manifold: long Newton paths amplify 1-ulp differences
    edge = compare(out, ref, TOL, TOL, TOL, 0.01)
    print(f"toy solver {solver}: edge chains {edge:.5f}, failed {(ref['status'] != 0).mean():.4f}")


@pytest.mark.parametrize("solver", [co.SOLVER_QUASI_NEWTON, co.SOLVER_NEWTON])
def test_fitzhugh_nagumo_16384_chains_match_oracle_per_chain(solver):
    """BASELINE config 4 at full size, one step of 0.01 (the C++ oracle needs ~3 s)."""
    import manifold_lifting_b200 as mlb

    g, y_obs = fn_data()
    cm = co.OracleModel.fn(y_obs, g["t_seq"], g["dt"], "unbounded")
    gm = mlb.models.FitzHughNagumoModel(y_obs, g["t_seq"], g["dt"], "unbounded")
    gs = mlb.IndependentAdditiveNoiseModelSystem(gm, gm.data, gm.dim_u)
    co.use_all_cores()
    q = fn_states(cm, 16384, seed=11)
    p = tangent_momentum(cm, q)
    dt = 0.01
    ref = cm.leapfrog_steps(q, p, dt, 1, co.solver_opts(solver))
    out = integrator(mlb, gs, dt, solver).steps(mlb.states.ChainState(q, p), 1)
    # momenta carry cond(C) * eps ~ 1e-6 of the vector (tests/test_cpu_ref_pinning.py)
    edge = compare(out, ref, 1e-9, 1e-6, 1e-9, 0.05, err=scaled_err)
    print(f"FN solver {solver}: edge chains {edge:.5f}")


def test_hodgkin_huxley_1024_chains_match_oracle_per_chain():
    """BASELINE config 5 per-GPU size (1,024 chains), Newton, one step of 0.002."""
    import manifold_lifting_b200 as mlb

    g, y_obs = hh_data()
    cm = co.OracleModel.hh(y_obs, g, "normal")
    gm = mlb.models.HodgkinHuxleyModel(y_obs, g, "normal")
    gs = mlb.IndependentAdditiveNoiseModelSystem(gm, gm.data, gm.dim_u)
    co.use_all_cores()
    q = hh_states(cm, g, 1024, seed=13)
    p = tangent_momentum(cm, q)
    dt = 0.002
    ref = cm.leapfrog_steps(q, p, dt, 1, co.solver_opts(co.SOLVER_NEWTON))
    out = integrator(mlb, gs, dt, co.SOLVER_NEWTON).steps(mlb.states.ChainState(q, p), 1)
    # residual noise floor ~1e-10..1e-9 against constraint_tol 1e-9: iteration counts of a
    # sizeable minority of solves move by one (see tests/test_cpu_ref_pinning.py); momenta
    # carry cond(C) * eps ~ 1e-4 of the vector
    st, ref_st = cpu(out.status), ref["status"]
    assert (st == ref_st).mean() > 0.97
    d_it = np.abs(cpu(out.iters_fwd) - ref["iters_fwd"]) + np.abs(cpu(out.iters_rev) - ref["iters_rev"])
    ok = (st == 0) & (ref_st == 0)
    assert ok.mean() > 0.6
    print("HH |d iters| histogram:", np.bincount(d_it[ok]))
    assert (d_it[ok] <= 2).mean() > 0.97 and (d_it[ok] == 0).mean() > 0.5
    assert scaled_err(cpu(out.pos)[ok], ref["q"][ok]) < 1e-8
    assert scaled_err(cpu(out.mom)[ok], ref["p"][ok]) < 1e-2
    print(f"HH: statuses equal {(st == ref_st).mean():.4f}, same counts {(d_it[ok] == 0).mean():.4f}, "
          f"q err {scaled_err(cpu(out.pos)[ok], ref['q'][ok]):.2e}, "
          f"p err {scaled_err(cpu(out.mom)[ok], ref['p'][ok]):.2e}")
