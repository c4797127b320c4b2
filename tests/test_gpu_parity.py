"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): per-step projected positions, momenta, Newton iteration
counts and accept/reject decisions within 1e-10 relative in fp64.
"""
import numpy as np
import pytest
import torch

from _cases import (EIGHT_SCHOOLS_SIGMA, EIGHT_SCHOOLS_Y, ChainState as OChainState,
                    hier_case, hier_states, rel_err, synthetic_schools, tangent_momentum,
                    toy_case, toy_states)  # fmt: skip
import c_oracle as co

pytestmark = pytest.mark.gpu
TOL = 1e-10


def same_iters(got, want, max_edge_frac=0.005):
    """Iteration counts must be EQUAL, except for chains that sit on a convergence
    threshold (|c| or |dq| within rounding of the tolerance at the deciding iteration),
    where a 1-ulp difference between the CPU and GPU arithmetic (FMA contraction,
    Newton-refined reciprocals) moves the exit by one iteration.  Those "edge" chains
    must be rare; they are returned so that the caller compares their positions at
    solver tolerance instead of 1e-10."""
    got, want = np.asarray(got), np.asarray(want)
    edge = got != want
    assert edge.mean() <= max_edge_frac, f"{edge.sum()} of {edge.size} iteration counts differ"
    return ~edge


def _mlb():
    import manifold_lifting_b200 as mlb

    return mlb


def gpu_hier(par="unbounded", y=EIGHT_SCHOOLS_Y, sigma=EIGHT_SCHOOLS_SIGMA):
    mlb = _mlb()
    model = mlb.models.EightSchoolsModel(y, sigma, par)
    return model, mlb.HierarchicalLatentVariableModelSystem(model, model.data, model.dim_u)


def gpu_toy(sigma=0.1, y=1.0, coeff=2.0):
    mlb = _mlb()
    model = mlb.models.Toy2DModel(sigma, y, coeff)
    return model, mlb.IndependentAdditiveNoiseModelSystem(model, model.data, model.dim_u)


def cpu(t):
    return t.detach().cpu().numpy()


CASES = [("hier", "unbounded"), ("hier", "normal"), ("toy", None)]


def make_case(kind, par, n, seed=7):
    if kind == "hier":
        omodel, osys, cm = hier_case(par)
        q = hier_states(omodel, n, seed)
        gm, gs = gpu_hier(par)
    else:
        omodel, osys, cm = toy_case(0.1)
        q = toy_states(omodel, n, seed)
        gm, gs = gpu_toy(0.1)
    return omodel, osys, cm, gm, gs, q


@pytest.mark.parametrize("kind,par", CASES)
def test_individual_ops_match_oracle(kind, par):
    """Every row of SURVEY 2b: constr, Jacobian blocks, Gram decomposition, log det and
    its gradient, J v, J^T w, pinv, inverse Jacobian product, normal-space component."""
    _, _, cm, gm, gs, q = make_case(kind, par, 257)
    rng = np.random.default_rng(3)
    n = q.shape[0]
    c = gs._constr(q)
    assert rel_err(cpu(c), cm.constr(q)) < TOL
    jac, c2 = gs._jacob_constr_blocks(q)
    ojac, oc = cm.jacob_constr_blocks(q)
    assert rel_err(cpu(c2), oc) < TOL
    for a, b in zip(jac, [x for x in ojac if x is not None]):
        assert rel_err(cpu(a), b) < TOL
    gram = gs._decompose_gram(jac)
    ogram = cm.decompose_gram(ojac)
    assert rel_err(cpu(gram[0]), ogram[0]) < TOL
    if len(gram) > 1:
        assert rel_err(cpu(gram[1]), ogram[1]) < TOL
    (val, aux), grad = gs._val_and_grad_log_det_sqrt_gram(q)
    og, oval = cm.grad_log_det_sqrt_gram(q)
    assert rel_err(cpu(val), oval) < TOL and rel_err(cpu(grad), og) < TOL
    val2, _ = gs._log_det_sqrt_gram(q)
    assert rel_err(cpu(val2), oval) < TOL
    v = rng.standard_normal((n, cm.Q))
    w = rng.standard_normal((n, cm.Y))
    assert rel_err(cpu(gs._lmult_by_jacob_constr(jac, v)), cm.lmult_by_jacob_constr(ojac, v)) < TOL
    assert rel_err(cpu(gs._rmult_by_jacob_constr(jac, w)), cm.rmult_by_jacob_constr(ojac, w)) < TOL
    assert rel_err(cpu(gs._lmult_by_pinv_jacob_constr(jac, gram, w)),
                   cm.lmult_by_pinv_jacob_constr(ojac, ogram, w)) < TOL  # fmt: skip
    assert rel_err(cpu(gs._normal_space_component(jac, gram, v)),
                   cm.normal_space_component(ojac, ogram, v)) < TOL  # fmt: skip
    q2 = q + 0.01 * rng.standard_normal(q.shape)
    jac2, _ = gs._jacob_constr_blocks(q2)
    ojac2, _ = cm.jacob_constr_blocks(q2)
    assert rel_err(cpu(gs._lmult_by_inv_jacob_product(jac2, jac, w)),
                   cm.lmult_by_inv_jacob_product(ojac2, ojac, w)) < TOL  # fmt: skip
    # reference-style call with unpacked blocks (systems.py:357-361)
    assert rel_err(cpu(gs._lmult_by_jacob_constr(*jac, v)), cm.lmult_by_jacob_constr(ojac, v)) < TOL
    # neg log dens + gradient, h1 / dh1_dpos
    mlb = _mlb()
    st = mlb.states.ChainState(q)
    oval, ograd = cm.neg_log_dens(q)
    assert rel_err(cpu(gs.neg_log_dens(st)), oval) < TOL
    assert rel_err(cpu(gs.grad_neg_log_dens(st)), ograd) < TOL
    assert rel_err(cpu(gs.dh1_dpos(st)), ograd + og) < TOL


SOLVERS = [co.SOLVER_QUASI_NEWTON, co.SOLVER_NEWTON, co.SOLVER_NEWTON_LS,
           co.SOLVER_MICI_NEWTON, co.SOLVER_MICI_QUASI_NEWTON]  # fmt: skip


@pytest.mark.parametrize("kind,par", CASES)
@pytest.mark.parametrize("solver", SOLVERS)
def test_projection_solvers_match_oracle(kind, par, solver):
    """Projected position, mu/dt, iteration counts, |dq|, |c|, status per chain."""
    _, _, cm, gm, gs, q = make_case(kind, par, 513)
    p = tangent_momentum(cm, q)
    dt = 0.2
    q_try = q + dt * p
    opts = co.solver_opts(solver)
    ref = cm.solve_projection(q_try, q, dt, opts)
    jac, _ = gs._jacob_constr_blocks(q)
    gram = gs._decompose_gram(jac)
    q_out, mu, it, ndq, err, n_c, status = gs._solve(
        solver, q_try, jac, gram, dt, 1e-9, 1e-8, 1e10, 50, "maximum")
    # Backtracking line search decides on |c(q + a d)| > |c(q)|: where such a comparison is
    # decided by rounding (both norms residual noise, or equal to 1e-6 relative - the oracle
    # flags those chains) the halving count and, on long wandering Newton paths (non-convex
    # toy constraint), the iteration path are not reproducible between two fp64
    # implementations.  Parity of status and iteration count is asserted on EVERY chain
    # outside a stated fraction of 1 % (measured: none of 513 eight-schools chains, 2 of 513
    # toy chains; the count carrying the oracle's flag is printed).
    reg = np.ones(len(q), bool)
    if solver == co.SOLVER_NEWTON_LS:
        amb = cm.last_line_search_ambiguous(len(q))
        mism = (cpu(status) != ref["status"]) | (np.abs(cpu(it) - ref["iters"]) > 1)
        print(f"{kind} {par} line search: {mism.sum()} of {len(q)} chains take another path, "
              f"{(mism & ~amb).sum()} of them unflagged; flagged overall {amb.mean():.3f}")
        assert mism.mean() <= 0.01
        reg = ~mism
    assert (cpu(status) == ref["status"])[reg].all()
    same = same_iters(cpu(it)[reg], ref["iters"][reg])
    assert np.abs(cpu(it)[reg] - ref["iters"][reg]).max() <= 1
    reg = reg.copy()
    reg[np.nonzero(reg)[0][~same]] = False
    assert (cpu(status) == ref["status"]).mean() > 0.99
    ok = (ref["status"] == 0) & reg
    assert ok.mean() > 0.8
    assert rel_err(cpu(q_out)[ok], ref["q"][ok]) < TOL
    assert rel_err(cpu(mu)[ok], ref["mu"][ok]) < TOL
    assert rel_err(cpu(err)[ok], ref["err"][ok]) < 1e-6  # |c| ~ 1e-12: rounding noise
    if solver == co.SOLVER_NEWTON_LS:
        # backtracking compares |c(q + a d)| > |c(q)|; once both are rounding noise
        # (<1e-12) the comparison, hence the halving count, is not reproducible
        same = cpu(n_c) == ref["n_constr"]
        assert same.mean() > 0.9 and (same | (ref["err"] < 1e-11))[ok].all()
    # invariant: converged chains are on the manifold
    assert np.abs(cpu(gs._constr(q_out))[ok]).max() < 1e-8


@pytest.mark.parametrize("kind,par", CASES)
@pytest.mark.parametrize("solver", SOLVERS)
def test_fused_leapfrog_steps_match_oracle(kind, par, solver):
    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case(kind, par, 1025)
    p = tangent_momentum(cm, q)
    dt, n_step = 0.15, 3
    ref = cm.leapfrog_steps(q, p, dt, n_step, co.solver_opts(solver))
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, dt, projection_solver={v: k for k, v in mlb.solvers.SOLVER_IDS.items()}[solver],
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    st = mlb.states.ChainState(q, p)
    out = integ.steps(st, n_step)
    reg = np.ones(len(q), bool)
    if solver == co.SOLVER_NEWTON_LS:  # see test_projection_solvers_match_oracle
        amb = cm.last_line_search_ambiguous(len(q))
        mism = ((cpu(out.status) != ref["status"]) | (cpu(out.n_done) != ref["n_done"])
                | (np.abs(cpu(out.iters_fwd) - ref["iters_fwd"]) > n_step)
                | (np.abs(cpu(out.iters_rev) - ref["iters_rev"]) > n_step))  # fmt: skip
        print(f"{kind} {par} line search: {mism.sum()} of {len(q)} chains take another path, "
              f"{(mism & ~amb).sum()} of them unflagged")
        assert mism.mean() <= 0.01
        reg = ~mism
    assert (cpu(out.status) == ref["status"])[reg].all()
    assert (cpu(out.n_done) == ref["n_done"])[reg].all()
    same = (same_iters(cpu(out.iters_fwd)[reg], ref["iters_fwd"][reg])
            & same_iters(cpu(out.iters_rev)[reg], ref["iters_rev"][reg]))
    edge = np.zeros(len(q), bool)
    edge[np.nonzero(reg)[0][~same]] = True
    if edge.any():  # threshold ties: same point to solver tolerance
        assert np.abs(cpu(out.pos)[edge] - ref["q"][edge]).max() < 1e-7
    reg = reg & ~edge
    assert (cpu(out.status) == ref["status"]).mean() > 0.99
    ok = (ref["status"] == 0) & reg
    assert ok.mean() > 0.8
    assert rel_err(cpu(out.pos)[ok], ref["q"][ok]) < TOL
    assert rel_err(cpu(out.mom)[ok], ref["p"][ok]) < TOL
    assert rel_err(cpu(out.h_init), ref["h_init"]) < TOL
    assert rel_err(cpu(out.h_final)[ok], ref["h_final"][ok]) < TOL
    # failed chains stay at their last completed step in both implementations
    bad = (ref["status"] != 0) & reg
    if bad.any():
        assert rel_err(cpu(out.pos)[bad], ref["q"][bad]) < TOL


@pytest.mark.parametrize("kind,par", CASES)
@pytest.mark.parametrize("n,n_step", [(1025, 3), (40000, 2)])
def test_steps_without_energy_outputs_are_bitwise_identical(kind, par, n, n_step):
    """`h_init` / `h_final` not requested (the bench's call): the kernels skip the VALUES of
    log det / prior density at interior points; positions, momenta, statuses and iteration
    counts must not move by a bit.  40,000 chains take the shared-memory-parked kernel form
    for the hierarchical model, 1,025 the register-resident one."""
    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case(kind, par, n)
    p = tangent_momentum(cm, q)
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, 0.15, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    a = integ.steps(mlb.states.ChainState(q, p), n_step)
    b = integ.steps(mlb.states.ChainState(q, p), n_step, energies=False)
    assert b.h_init is None and b.h_final is None
    assert torch.equal(a.pos, b.pos) and torch.equal(a.mom, b.mom)
    assert torch.equal(a.status, b.status) and torch.equal(a.n_done, b.n_done)
    assert torch.equal(a.iters_fwd, b.iters_fwd) and torch.equal(a.iters_rev, b.iters_rev)
    ok = cpu(a.status) == 0
    ref = cm.leapfrog_steps(q, p, 0.15, n_step, co.solver_opts(co.SOLVER_NEWTON))
    both = ok & (ref["status"] == 0)
    assert rel_err(cpu(a.h_final)[both], ref["h_final"][both]) < TOL
    assert rel_err(cpu(a.h_init), ref["h_init"]) < TOL


@pytest.mark.parametrize("kind,par", [("hier", "unbounded"), ("toy", None)])
def test_unfused_step_equals_fused_step(kind, par):
    """Driving the system method by method (as Mici does) gives the fused kernel's
    result: the boundary is a true drop-in."""
    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case(kind, par, 64)
    p = tangent_momentum(cm, q)
    ref = cm.leapfrog_steps(q, p, 0.1, 1, co.solver_opts(co.SOLVER_NEWTON))
    ok = ref["status"] == 0
    q, p = q[ok], p[ok]
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, 0.1, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    st = mlb.states.ChainState(q, p)
    a, b = integ.steps(st, 1), integ.step_unfused(st)
    assert rel_err(cpu(a.pos), cpu(b.pos)) < TOL and rel_err(cpu(a.mom), cpu(b.mom)) < TOL
    assert (cpu(a.iters_fwd) == cpu(b.iters_fwd)).all()


def test_step_matches_autodiff_oracle_single_chains():
    """Against the line-by-line torch restatement (autodiff Jacobians, reference loops)."""
    from mlift_oracle import solvers as osolvers
    from mlift_oracle.mici import ConstrainedLeapfrogIntegrator as OInteg

    mlb = _mlb()
    omodel, osys, cm, gm, gs, q = make_case("hier", "unbounded", 6)
    p = tangent_momentum(cm, q)
    for sfn, gfn in [
        (osolvers.solve_projection_onto_manifold_newton,
         mlb.jitted_solve_projection_onto_manifold_newton),
        (osolvers.solve_projection_onto_manifold_quasi_newton,
         mlb.jitted_solve_projection_onto_manifold_quasi_newton),
        (osolvers.solve_projection_onto_manifold_newton_with_line_search,
         mlb.jitted_solve_projection_onto_manifold_newton_with_line_search),
    ]:  # fmt: skip
        kw = dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50)
        out = mlb.integrators.ConstrainedLeapfrogIntegrator(
            gs, 0.15, projection_solver=gfn, projection_solver_kwargs=kw
        ).steps(mlb.states.ChainState(q, p), 1)
        for i in range(q.shape[0]):
            integ = OInteg(osys, 0.15, projection_solver=sfn, projection_solver_kwargs=kw)
            s1 = integ.step(OChainState(q[i], p[i]))
            assert rel_err(cpu(out.pos)[i], s1.pos) < TOL
            assert rel_err(cpu(out.mom)[i], s1.mom) < TOL
            assert (int(out.iters_fwd[i]), int(out.iters_rev[i])) == integ.last_iters


@pytest.mark.parametrize("kind,par", [("hier", "unbounded"), ("hier", "normal"), ("toy", None)])
def test_hmc_transitions_match_oracle_decisions(kind, par):
    """Accept / reject decisions, accept_stat and new positions with supplied noise."""
    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case(kind, par, 777)
    rng = np.random.default_rng(5)
    n, Q = q.shape
    n_iter, n_step, dt = 3, 4, 0.2
    noise = rng.standard_normal((n_iter, n, Q))
    unif = rng.uniform(size=(n_iter, n))
    opts = co.solver_opts(co.SOLVER_NEWTON)
    qo = q.copy()
    acc_ref = []
    for it in range(n_iter):
        r = cm.hmc_transition(qo, dt, n_step, opts, mom_noise=noise[it], unif=unif[it])
        qo = r["q"]
        acc_ref.append(r["accept_stat"])
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, dt, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    sampler = mlb.samplers.StaticMetropolisHMC(gs, integ, n_step=n_step, seed=1)
    qg = mlb._lib.to_soa(q)
    qg = mlb._lib.soa_clone(qg)
    tr_acc = torch.empty(n_iter, n, dtype=torch.float64, device="cuda")
    noise_dev = torch.as_tensor(noise.transpose(0, 2, 1).copy(), device="cuda")
    sampler._launch(qg, n_iter, trace_accept=tr_acc, mom_noise=noise_dev,
                    unif=torch.as_tensor(unif, device="cuda"))  # fmt: skip
    assert rel_err(cpu(tr_acc), np.stack(acc_ref)) < 1e-9
    assert rel_err(cpu(qg), qo) < TOL


def test_philox_stream_matches_oracle():
    L = _mlb()._lib.lib()
    for chain in (0, 5, 2**33 + 17):
        out = np.empty(18)
        L.mlb_philox_normals_host(1234, chain, 7, 18, out.ctypes.data)
        assert rel_err(out, co.philox_normals(1234, chain, 7, 18)) < 1e-14
        assert abs(L.mlb_philox_uniform_host(1234, chain, 7) - co.philox_uniform(1234, chain, 7)) < 1e-16


def test_device_philox_transitions_match_oracle():
    """Device-generated momentum / uniforms follow the same Philox streams as the oracle,
    so whole sampled trajectories agree (sharding-independent RNG)."""
    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 300)
    opts = co.solver_opts(co.SOLVER_NEWTON)
    qo = q.copy()
    for it in range(2):
        qo = cm.hmc_transition(qo, 0.2, 3, opts, seed=99, chain0=1000, iteration=it)["q"]
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, 0.2, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    sampler = mlb.samplers.StaticMetropolisHMC(gs, integ, n_step=3, seed=99)
    qg = mlb._lib.soa_clone(mlb._lib.to_soa(q))
    sampler._launch(qg, 2, chain_offset=1000)
    assert rel_err(cpu(qg), qo) < 1e-9


def test_host_buffer_entry_point_matches_device_path():
    import ctypes as C

    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 1000)
    p = tangent_momentum(cm, q)
    ref = cm.leapfrog_steps(q, p, 0.1, 2, co.solver_opts(co.SOLVER_NEWTON))
    qh, ph = q.copy(), p.copy()
    status = np.empty(q.shape[0], np.int32)
    opts = mlb._lib.SolverOpts(mlb._lib.SOLVER_NEWTON, 50, 10, 0, 1e-9, 1e-8, 1e10, 2e-8, 0, 0)
    rc = mlb._lib.lib().mlb_leapfrog_steps_host(
        gm.handle, C.c_int64(q.shape[0]), qh.ctypes.data_as(C.c_void_p),
        ph.ctypes.data_as(C.c_void_p), C.c_double(0.1), C.c_int32(2), C.byref(opts),
        status.ctypes.data_as(C.c_void_p), None, None, None, None, None)
    assert rc == 0
    assert (status == ref["status"]).all()
    assert rel_err(qh, ref["q"]) < TOL and rel_err(ph, ref["p"]) < TOL


def test_host_buffer_entry_point_chunk_pipeline():
    """More chains than one chunk of the host entry point's copy / compute pipeline (16,384
    chains per chunk, ragged tail): every output array, against the device-pointer path."""
    import ctypes as C

    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case("hier", "unbounded", 40000)
    p = tangent_momentum(cm, q)
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, 0.1, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    dev = integ.steps(mlb.states.ChainState(q, p), 3, raise_on_failure=False)
    n = q.shape[0]
    qh, ph = q.copy(), p.copy()
    i32 = [np.full(n, -7, np.int32) for _ in range(4)]
    f64 = [np.full(n, np.nan) for _ in range(2)]
    opts = mlb._lib.SolverOpts(mlb._lib.SOLVER_NEWTON, 50, 10, 0, 1e-9, 1e-8, 1e10, 2e-8, 0, 0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for _ in range(2):  # second call reuses the staging area and the streams
        qh[:], ph[:] = q, p
        rc = mlb._lib.lib().mlb_leapfrog_steps_host(
            gm.handle, C.c_int64(n), vp(qh), vp(ph), C.c_double(0.1), C.c_int32(3),
            C.byref(opts), vp(i32[0]), vp(i32[1]), vp(i32[2]), vp(i32[3]), vp(f64[0]), vp(f64[1]))
        assert rc == 0
        assert (i32[0] == cpu(dev.status)).all() and (i32[1] == cpu(dev.n_done)).all()
        assert (i32[2] == cpu(dev.iters_fwd)).all() and (i32[3] == cpu(dev.iters_rev)).all()
        assert np.array_equal(qh, cpu(dev.pos)) and np.array_equal(ph, cpu(dev.mom))
        # positions / momenta / counts are bit-identical; the Hamiltonians come from two
        # kernels (the 40,000-chain device call takes the shared-memory-parked form, the
        # 16,384-chain chunks the register-resident one) whose compilers fuse the final
        # nld + log det + |p|^2 / 2 additions differently: last-bit agreement
        assert np.allclose(f64[0], cpu(dev.h_init), rtol=1e-14, atol=0)
        assert np.allclose(f64[1], cpu(dev.h_final), rtol=1e-14, atol=0, equal_nan=True)


@pytest.mark.parametrize("kind,solver,dt,max_iters,rev_tol", [
    ("hier", "newton", 1.0, 8, 2e-8), ("hier", "quasi_newton", 0.4, 9, 2e-8),
    ("hier", "newton_line_search", 1.0, 8, 2e-8), ("toy", "newton", 0.15, 50, 2e-8),
    ("toy", "quasi_newton", 0.15, 50, 2e-8)])
def test_small_launch_form_matches_large_batch(kind, solver, dt, max_iters, rev_tol):
    """Launches below ~16,000 chains take the two-warp form of the fused trajectory kernel
    (csrc/mlb_duo.cuh: reversibility check on a second warp, one step behind, speculative
    steps discarded on a failed verdict); larger ones the one-warp forms.  The same chains
    run alone (small launch) and as the head of a 40,000-chain batch must give the same
    statuses, completed steps, iteration counts and states (to rounding: the candidate
    momentum of a step is formed differently), failures included: step sizes and iteration
    caps make eight-schools chains fail at different steps in the forward and in the reverse
    solve, and the toy model's chains also fail the reversibility check itself (for 3,000
    chains x 6 steps at dt = 0.15 the oracle counts 104 not converged, 57 non-reversible)."""
    mlb = _mlb()
    _, _, cm, gm, gs, q = make_case(kind, "unbounded" if kind == "hier" else None, 40000)
    p = tangent_momentum(cm, q)
    fn = {"newton": mlb.jitted_solve_projection_onto_manifold_newton,
          "quasi_newton": mlb.jitted_solve_projection_onto_manifold_quasi_newton,
          "newton_line_search":
              mlb.jitted_solve_projection_onto_manifold_newton_with_line_search}[solver]
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, dt, projection_solver=fn, reverse_check_tol=rev_tol,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8,
                                      max_iters=max_iters))
    big = integ.steps(mlb.states.ChainState(q, p), 6, raise_on_failure=False)
    m = 3000
    small = integ.steps(mlb.states.ChainState(q[:m], p[:m]), 6, raise_on_failure=False)
    st_b, st_s = cpu(big.status)[:m], cpu(small.status)
    nd_b, nd_s = cpu(big.n_done)[:m], cpu(small.n_done)
    assert (st_s != 0).any() and (st_s == 0).any()  # both outcomes are exercised
    if kind == "toy" and solver == "newton":
        assert (st_s == 3).sum() > 10 and (st_s == 2).sum() > 10  # failed checks, failed solves
    # a solve decided by rounding (iteration count tie at a tolerance) may differ: rare
    same = (st_b == st_s) & (nd_b == nd_s) & (cpu(big.iters_fwd)[:m] == cpu(small.iters_fwd)) \
        & (cpu(big.iters_rev)[:m] == cpu(small.iters_rev))
    assert same.mean() > 0.995, same.mean()
    assert rel_err(cpu(small.pos)[same], cpu(big.pos)[:m][same]) < 1e-10
    assert rel_err(cpu(small.mom)[same], cpu(big.mom)[:m][same]) < 1e-10
    ok = same & (st_s == 0)
    assert np.allclose(cpu(small.h_final)[ok], cpu(big.h_final)[:m][ok], rtol=1e-11, atol=0)
    assert np.allclose(cpu(small.h_init), cpu(big.h_init)[:m], rtol=1e-13, atol=0)


def test_full_size_properties_65536_chains():
    """BASELINE config 3 at full size: size-independent properties (the oracle is not run
    at this size): on-manifold, tangent momentum, energy error O(dt^2), reversibility."""
    mlb = _mlb()
    y, sigma = synthetic_schools(8)
    gm, gs = gpu_hier("unbounded", y, sigma)
    rng = np.random.default_rng(202101)
    n = 65536
    q = gm.sample_initial_states(rng, n)
    q[:, 0] = np.clip(q[:, 0], -10, 10)
    q[:, 1] = np.clip(q[:, 1], -2, 3)
    par = gm.generate_params(q[:, :2])
    q[:, 10:] = (y[None] - par["μ"][:, None] - par["τ"][:, None] * q[:, 2:10]) / sigma[None]
    st = mlb.states.ChainState(q)
    st.mom = gs.sample_momentum(st, torch.Generator(device="cuda").manual_seed(1))
    integ = mlb.integrators.ConstrainedLeapfrogIntegrator(
        gs, 0.05, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    out = integ.steps(st, 8)
    ok = out.status == 0
    assert float(ok.double().mean()) > 0.99
    assert float(gs._constr(out.pos)[ok].abs().max()) < 1e-8
    jp = gs.lmult_by_jacob_constr(mlb.states.ChainState(out.pos), out.mom)
    assert float(jp[ok].abs().max()) < 1e-9
    dh = (out.h_final - out.h_init)[ok].abs()
    assert float(dh.median()) < 0.05
    # reversibility: flip direction, integrate back
    back = mlb.states.ChainState(out.pos[ok], out.mom[ok], dir=-1)
    back = integ.steps(back, 8)
    ok2 = back.status == 0
    assert float((back.pos[ok2] - st.pos[ok][ok2]).abs().max()) < 1e-6


def test_fast_math_primitives():
    """Newton-refined MUFU reciprocal / sqrt / rsqrt (csrc/mlb_common.cuh) vs the IEEE
    operations: <= 2 ulp over 60 decades."""
    import ctypes as C

    mlb = _mlb()
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, 200000), rng.uniform(0.5, 2.0, 100000),
                        [1.0, 2.0, 0.25, 3.0, 1e-300, 1e300]])  # fmt: skip
    xd = torch.as_tensor(x, device="cuda")
    r, s, rs = (torch.empty_like(xd) for _ in range(3))
    mlb._lib.check(mlb._lib.lib().mlb_selftest_fast_math(
        C.c_int64(x.size), C.c_void_p(xd.data_ptr()), C.c_void_p(r.data_ptr()),
        C.c_void_p(s.data_ptr()), C.c_void_p(rs.data_ptr()), mlb._lib.stream_ptr()))
    eps = np.finfo(np.float64).eps
    assert np.abs(cpu(r) * x - 1.0).max() < 2 * eps
    assert np.abs(cpu(s) / np.sqrt(x) - 1.0).max() < 2 * eps
    assert np.abs(cpu(rs) * np.sqrt(x) - 1.0).max() < 2 * eps
    neg = cpu(-r)  # sign symmetry
    mlb._lib.check(mlb._lib.lib().mlb_selftest_fast_math(
        C.c_int64(x.size), C.c_void_p((-xd).data_ptr()), C.c_void_p(r.data_ptr()),
        C.c_void_p(s.data_ptr()), C.c_void_p(rs.data_ptr()), mlb._lib.stream_ptr()))
    assert np.array_equal(cpu(r), neg)


def test_branch_free_exp_expm1():
    """Straight-line exp / expm1 used by the Hodgkin-Huxley recurrence
    (csrc/mlb_common.cuh) against mpmath-grade references (numpy long double where
    available, else float64 libm): <= 2 ulp (exp), <= 3 ulp (expm1) on [-700, 700], dense
    sampling around 0 where expm1 cancels, NaN propagated."""
    import ctypes as C

    mlb = _mlb()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-40, 40, 200000),
                        rng.uniform(-1, 1, 200000), 10.0 ** rng.uniform(-300, 0, 50000),
                        -(10.0 ** rng.uniform(-300, 0, 50000)),
                        [0.0, -0.0, 0.5 * np.log(2), -0.5 * np.log(2), 700.0, -700.0]])  # fmt: skip
    xd = torch.as_tensor(x, device="cuda")
    e, em1 = torch.empty_like(xd), torch.empty_like(xd)
    mlb._lib.check(mlb._lib.lib().mlb_selftest_exp(
        C.c_int64(x.size), C.c_void_p(xd.data_ptr()), C.c_void_p(e.data_ptr()),
        C.c_void_p(em1.data_ptr()), mlb._lib.stream_ptr()))
    xl = x.astype(np.longdouble)
    ref_e, ref_m = np.exp(xl), np.expm1(xl)
    eps = np.finfo(np.float64).eps
    assert np.abs((cpu(e).astype(np.longdouble) - ref_e) / ref_e).max() < 2 * eps
    nz = x != 0.0
    assert np.abs((cpu(em1)[nz].astype(np.longdouble) - ref_m[nz]) / ref_m[nz]).max() < 3 * eps
    assert (cpu(em1)[~nz] == 0.0).all()
    bad = torch.tensor([float("nan")], dtype=torch.float64, device="cuda")
    mlb._lib.check(mlb._lib.lib().mlb_selftest_exp(
        C.c_int64(1), C.c_void_p(bad.data_ptr()), C.c_void_p(e.data_ptr()),
        C.c_void_p(em1.data_ptr()), mlb._lib.stream_ptr()))
    assert np.isnan(cpu(e)[0]) and np.isnan(cpu(em1)[0])

