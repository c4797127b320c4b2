"""Edge cases of the fused step through the C ABI, against the oracle: ragged and tiny
batches (the kernels work on 32- / 64-chain blocks), padded leading dimension, n_step = 0
(only h_init is produced), every failure status (the reference raises ConvergenceError /
NonReversibleStepError per chain, solvers.py:86-95; here they are
per-chain codes and the chain is handed back at its last completed step), per-chain step
sizes, and the single-chain exception behaviour of the reference-shaped API."""
import numpy as np
import pytest
import torch

import c_oracle as co
from _cases import fn_data, fn_states, rel_err, tangent_momentum
from test_gpu_parity import TOL, cpu, make_case, same_iters

pytestmark = pytest.mark.gpu


def _integ(mlb, gs, dt, solver=co.SOLVER_NEWTON, **kw):
    kwargs = dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50)
    kwargs.update(kw)
    return mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, dt, projection_solver={v: k for k, v in mlb.solvers.SOLVER_IDS.items()}[solver],
        projection_solver_kwargs=kwargs)


@pytest.mark.parametrize("n", [1, 2, 31, 33, 63, 65, 127, 1000])
def test_ragged_batch_sizes_hier(n):
    import manifold_lifting_b200 as mlb

    _, _, cm, gm, gs, q = make_case("hier", "unbounded", n, seed=n)
    p = tangent_momentum(cm, q)
    ref = cm.leapfrog_steps(q, p, 0.1, 2, co.solver_opts(co.SOLVER_NEWTON))
    out = _integ(mlb, gs, 0.1).steps(mlb.states.ChainState(q, p), 2, raise_on_failure=False)
    assert (cpu(out.status) == ref["status"]).all() and (cpu(out.n_done) == ref["n_done"]).all()
    ok = (ref["status"] == 0) & (cpu(out.iters_fwd) == ref["iters_fwd"]) \
        & (cpu(out.iters_rev) == ref["iters_rev"])
    assert ok.mean() >= 0.9 or n < 10
    if ok.any():
        assert rel_err(cpu(out.pos)[ok], ref["q"][ok]) < TOL
        assert rel_err(cpu(out.mom)[ok], ref["p"][ok]) < TOL


@pytest.mark.parametrize("n", [1, 5, 33])
def test_ragged_batch_sizes_fn_sub_phase_kernels(n):
    """The ODE family's chunked CTAs and role-split sweeps with partially filled warps."""
    import manifold_lifting_b200 as mlb

    g, y_obs = fn_data()
    cm = co.OracleModel.fn(y_obs, g["t_seq"], g["dt"], "unbounded")
    gm = mlb.models.FitzHughNagumoModel(y_obs, g["t_seq"], g["dt"], "unbounded")
    gs = mlb.IndependentAdditiveNoiseModelSystem(gm, gm.data, gm.dim_u)
    q = fn_states(cm, n)
    p = tangent_momentum(cm, q)
    ref = cm.leapfrog_steps(q, p, 0.01, 1, co.solver_opts(co.SOLVER_NEWTON))
    out = _integ(mlb, gs, 0.01).steps(mlb.states.ChainState(q, p), 1, raise_on_failure=False)
    assert (cpu(out.status) == ref["status"]).all()
    ok = (ref["status"] == 0) & same_iters(cpu(out.iters_fwd), ref["iters_fwd"], 1.0) \
        & same_iters(cpu(out.iters_rev), ref["iters_rev"], 1.0)
    if ok.any():
        scale = np.abs(ref["q"][ok]).max(1, keepdims=True)
        assert (np.abs(cpu(out.pos)[ok] - ref["q"][ok]) / np.maximum(scale, 1.0)).max() < 1e-9


def test_padded_leading_dimension_and_empty_batch():
    """ld > n (chains are columns of a wider allocation) and n = 0."""
    import ctypes as C

    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 100)
    p = tangent_momentum(cm, q)
    ref = cm.leapfrog_steps(q, p, 0.1, 2, co.solver_opts(co.SOLVER_NEWTON))
    n, Q, ld = 100, q.shape[1], 160
    qd = torch.full((Q, ld), float("nan"), dtype=torch.float64, device="cuda")
    pd = torch.full((Q, ld), float("nan"), dtype=torch.float64, device="cuda")
    qd[:, :n] = torch.as_tensor(q.T).cuda()
    pd[:, :n] = torch.as_tensor(p.T).cuda()
    status = torch.full((ld,), -7, dtype=torch.int32, device="cuda")
    opts = _lib.SolverOpts(_lib.SOLVER_NEWTON, 50, 10, 0, 1e-9, 1e-8, 1e10, 2e-8, 0, 0)
    L = _lib.lib()

    def call(n_chain):
        return L.mlb_leapfrog_steps(
            gm.handle, C.c_int64(n_chain), C.c_int64(ld), C.c_void_p(qd.data_ptr()),
            C.c_void_p(pd.data_ptr()), C.c_double(0.1), None, C.c_int32(2), C.byref(opts),
            C.c_void_p(status.data_ptr()), None, None, None, None, None, _lib.stream_ptr())

    assert call(0) == 0
    torch.cuda.synchronize()
    assert (status == -7).all()
    assert call(n) == 0
    torch.cuda.synchronize()
    assert (cpu(status[:n]) == ref["status"]).all() and (status[n:] == -7).all()
    ok = ref["status"] == 0
    assert rel_err(cpu(qd[:, :n]).T[ok], ref["q"][ok]) < TOL
    assert torch.isnan(qd[:, n:]).all() and torch.isnan(pd[:, n:]).all()  # padding untouched


def test_zero_steps_returns_initial_hamiltonian_only():
    import manifold_lifting_b200 as mlb

    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 200)
    p = tangent_momentum(cm, q)
    ref = cm.leapfrog_steps(q, p, 0.1, 0, co.solver_opts(co.SOLVER_NEWTON))
    out = _integ(mlb, gs, 0.1).steps(mlb.states.ChainState(q, p), 0, raise_on_failure=False)
    assert (cpu(out.status) == 0).all() and (cpu(out.n_done) == 0).all()
    assert np.array_equal(cpu(out.pos), q) and np.array_equal(cpu(out.mom), p)
    assert rel_err(cpu(out.h_init), ref["h_init"]) < TOL


def test_every_failure_status_matches_oracle():
    """NOT_CONVERGED (max_iters too small), DIVERGED (residual above divergence_tol),
    NON_REVERSIBLE (reverse_check_tol below the solver's accuracy): same code per chain as
    the oracle, and failed chains are left at their last completed step."""
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 512, seed=21)
    p = 3.0 * tangent_momentum(cm, q)  # energetic: first residuals are large
    seen = set()
    # (quasi-Newton retractions are accurate to ~1e-9, so a 1e-11 reverse check fails about
    # half of the chains; Newton's quadratic convergence would pass it)
    for solver, kw, rev_tol in ((co.SOLVER_NEWTON, dict(max_iters=2), 2e-8),
                                (co.SOLVER_NEWTON, dict(divergence_tol=1e-3), 2e-8),
                                (co.SOLVER_QUASI_NEWTON, dict(), 1e-11)):  # fmt: skip
        o = co.solver_opts(solver, max_iters=kw.get("max_iters", 50),
                           divergence_tol=kw.get("divergence_tol", 1e10),
                           reverse_check_tol=rev_tol)  # fmt: skip
        ref = cm.leapfrog_steps(q, p, 0.3, 3, o)
        integ = _integ(mlb, gs, 0.3, solver, **kw)
        integ.reverse_check_tol = rev_tol
        out = integ.steps(mlb.states.ChainState(q, p), 3, raise_on_failure=False)
        st = cpu(out.status)
        assert (st == ref["status"]).mean() > 0.97
        same = st == ref["status"]
        assert (cpu(out.n_done)[same] == ref["n_done"][same]).all()
        bad = same & (st != 0)
        assert bad.any()
        assert rel_err(cpu(out.pos)[bad], ref["q"][bad]) < 1e-9
        seen |= set(np.unique(st[bad]).tolist())
    assert {_lib.ST_DIVERGED, _lib.ST_NOT_CONVERGED, _lib.ST_NON_REVERSIBLE} <= seen


def test_per_chain_step_sizes_and_direction():
    import manifold_lifting_b200 as mlb

    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 300, seed=9)
    p = tangent_momentum(cm, q)
    rng = np.random.default_rng(0)
    dts = rng.uniform(0.02, 0.2, q.shape[0])
    ref = cm.leapfrog_steps(q, p, 0.0, 2, co.solver_opts(co.SOLVER_NEWTON), dt_per_chain=-dts)
    integ = _integ(mlb, gs, torch.as_tensor(dts).cuda())
    out = integ.steps(mlb.states.ChainState(q, p, dir=-1), 2, raise_on_failure=False)
    assert (cpu(out.status) == ref["status"]).all()
    ok = (ref["status"] == 0) & (cpu(out.iters_fwd) == ref["iters_fwd"])
    assert ok.mean() > 0.9
    assert rel_err(cpu(out.pos)[ok], ref["q"][ok]) < TOL
    assert rel_err(cpu(out.mom)[ok], ref["p"][ok]) < TOL


def test_single_chain_raises_like_the_reference():
    """One chain through the reference-shaped API: failures are Mici-style exceptions
    (solvers.py:86-95; ConstrainedLeapfrogIntegrator raises NonReversibleStepError)."""
    import manifold_lifting_b200 as mlb

    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 1, seed=2)
    p = 3.0 * tangent_momentum(cm, q)
    state = mlb.states.ChainState(q, p)
    with pytest.raises(mlb.errors.ConvergenceError):
        _integ(mlb, gs, 0.3, max_iters=1).step(state)
    integ = _integ(mlb, gs, 0.05)
    integ.reverse_check_tol = 1e-17
    with pytest.raises(mlb.errors.NonReversibleStepError):
        integ.step(mlb.states.ChainState(q, tangent_momentum(cm, q)))
    assert issubclass(mlb.errors.NonReversibleStepError, mlb.errors.IntegratorError)
    assert issubclass(mlb.errors.ConvergenceError, mlb.errors.IntegratorError)
