"""Batched "tridiagonal + low rank" Gram algebra of the state-space systems on the GPU
(csrc/mlb_ssm.cu, through the C ABI) against the NumPy restatement of the reference's
closures (oracle/mlift_oracle/state_space.py, mlift/systems.py:1161-1243) and, at full
size, through properties that do not need the oracle."""
import numpy as np
import pytest
import torch

from mlift_oracle import state_space as oss

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (2, 2), (3, 1), (9, 3), (40, 8), (200, 3), (1000, 4)]


def _batch(rng, n, dim_y, dim_u):
    chains = [oss.random_blocks(rng, dim_y, dim_u) for _ in range(n)]
    du = np.stack([c[0] for c in chains])
    dv = np.stack([c[1] for c in chains])
    dn = (np.stack([c[2][0] for c in chains]), np.stack([c[2][1] for c in chains]))
    dx = np.stack([c[3] for c in chains])
    return chains, (du, dv, dn, dx)


def _close(x, ref, tol):
    x = x.cpu().numpy() if isinstance(x, torch.Tensor) else x
    assert x.shape == ref.shape
    if ref.size == 0:
        return
    assert np.abs(x - ref).max() <= tol * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dim_y,dim_u", SHAPES)
def test_state_space_gram_ops_match_restatement(dim_y, dim_u):
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    rng = np.random.default_rng(1000 * dim_y + dim_u)
    n = 67 if dim_y < 1000 else 33
    chains, arrs = _batch(rng, n, dim_y, dim_u)
    chains_r, arrs_r = _batch(rng, n, dim_y, dim_u)
    J, Jr = ss.JacobBlocks(*arrs), ss.JacobBlocks(*arrs_r)
    vq = rng.normal(size=(n, dim_u + 2 * dim_y))
    vy = rng.normal(size=(n, dim_y))

    _close(ss.lmult_by_jacob_constr(J, None, None, None, vq),
           np.stack([oss.lmult_by_jacob_constr(*chains[i], vq[i]) for i in range(n)]), 1e-14)  # fmt: skip
    _close(ss.rmult_by_jacob_constr(J, None, None, None, vy),
           np.stack([oss.rmult_by_jacob_constr(*chains[i], vy[i]) for i in range(n)]), 1e-13)  # fmt: skip

    a, b, chol = ss.decompose_gram(J)
    ref = [oss.decompose_gram(*chains[i]) for i in range(n)]
    _close(a, np.stack([r[0] for r in ref]), 1e-15)
    _close(b, np.stack([r[1] for r in ref]), 1e-15)
    _close(chol, np.stack([r[2] for r in ref]), 1e-12)
    assert float(torch.triu(chol, 1).abs().max()) == 0.0

    _close(ss.lmult_by_inv_gram(J, None, None, None, a, b, chol, vy),
           np.stack([oss.lmult_by_inv_gram(*chains[i], *ref[i], vy[i]) for i in range(n)]),
           1e-12)  # fmt: skip
    # the raw-array form of the same call (blocks converted on the fly)
    _close(ss.lmult_by_inv_gram(*arrs, a, b, chol, vy),
           ss.lmult_by_inv_gram(J, None, None, None, a, b, chol, vy).cpu().numpy(), 0.0)
    _close(ss.lmult_by_inv_jacob_product(J, None, None, None, Jr, None, None, None, vy),
           np.stack([oss.lmult_by_inv_jacob_product(*chains[i], *chains_r[i], vy[i])
                     for i in range(n)]), 1e-11)  # fmt: skip
    _close(ss.log_det_sqrt_gram(J, gram=(a, b, chol)),
           np.array([oss.log_det_sqrt_gram(*chains[i]) for i in range(n)]), 1e-12)
    _close(ss.log_det_sqrt_gram(J), ss.log_det_sqrt_gram(J, gram=(a, b, chol)).cpu().numpy(), 0.0)


@pytest.mark.parametrize("scale", [1e-85, 1e85])
def test_state_space_log_det_with_badly_scaled_pivots(scale):
    """The kernel takes one logarithm per PAIR of pivots (and of |dx_dy| entries); with
    pivots around 1e-170 / 1e+170 the pair products leave the double range, where the
    reference's per-term recursion (linalg.py:104-114) stays finite.  The log-determinant
    must follow: log det sqrt Gram(s J) = log det sqrt Gram(J) + dim_y log s."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    rng = np.random.default_rng(8)
    n, dim_y, dim_u = 33, 37, 3
    chains, (du, dv, dn, dx) = _batch(rng, n, dim_y, dim_u)
    base = ss.log_det_sqrt_gram(ss.JacobBlocks(du, dv, dn, dx)).cpu().numpy()
    scaled = ss.log_det_sqrt_gram(
        ss.JacobBlocks(du * scale, dv * scale, (dn[0] * scale, dn[1] * scale), dx)).cpu().numpy()
    assert np.isfinite(scaled).all()
    assert np.abs(scaled - (base + dim_y * np.log(scale))).max() < 1e-9 * dim_y * abs(np.log(scale))
    ref = np.array([oss.log_det_sqrt_gram(c[0] * scale, c[1] * scale,
                                          (c[2][0] * scale, c[2][1] * scale), c[3])
                    for c in chains])
    assert np.abs(scaled - ref).max() < 1e-9 * np.abs(ref).max()
    # |dx_dy| around 1e+-170: pair products of the second sum leave the range as well
    dxs = ss.log_det_sqrt_gram(ss.JacobBlocks(du, dv, dn, dx * scale**2)).cpu().numpy()
    assert np.abs(dxs - (base - dim_y * np.log(scale**2))).max() < 1e-9 * dim_y * abs(np.log(scale**2))


def test_state_space_gram_ops_match_dense_linear_algebra():
    """Independent of the restated scans: J assembled densely, G = J J^T."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    rng = np.random.default_rng(7)
    n, dim_y, dim_u = 19, 30, 3
    chains, arrs = _batch(rng, n, dim_y, dim_u)
    chains_r, arrs_r = _batch(rng, n, dim_y, dim_u)
    J, Jr = ss.JacobBlocks(*arrs), ss.JacobBlocks(*arrs_r)
    vy = rng.normal(size=(n, dim_y))
    Jd = [oss.dense_jacobian(*c) for c in chains]
    Jrd = [oss.dense_jacobian(*c) for c in chains_r]
    a, b, chol = ss.decompose_gram(J)
    _close(ss.lmult_by_inv_gram(J, None, None, None, a, b, chol, vy),
           np.stack([np.linalg.solve(Jd[i] @ Jd[i].T, vy[i]) for i in range(n)]), 1e-11)
    _close(ss.lmult_by_inv_jacob_product(J, None, None, None, Jr, None, None, None, vy),
           np.stack([np.linalg.solve(Jd[i] @ Jrd[i].T, vy[i]) for i in range(n)]), 1e-10)
    _close(ss.log_det_sqrt_gram(J, gram=(a, b, chol)),
           np.array([0.5 * np.linalg.slogdet(Jd[i] @ Jd[i].T)[1] - np.log(np.abs(chains[i][3])).sum()
                     for i in range(n)]), 1e-11)  # fmt: skip


def test_state_space_failure_and_argument_checks():
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    ss = mlb.state_space
    rng = np.random.default_rng(3)
    _, arrs = _batch(rng, 5, 6, 2)
    with pytest.raises(_lib.MlbError):  # dim_u above MLB_SSM_MAX_DIM_U
        ss.decompose_gram(np.zeros((4, 3, 9)), np.ones((4, 3)), (np.ones((4, 3)), np.ones((4, 2))))
    with pytest.raises(_lib.MlbError):  # ragged block
        ss.JacobBlocks(arrs[0], arrs[1][:, :-1], arrs[2])
    with pytest.raises(_lib.MlbError):  # log-det needs dx_dy
        ss.log_det_sqrt_gram(arrs[0], arrs[1], arrs[2])
    # NaN blocks give a NaN factor for that chain only (no exception, no contamination)
    du = arrs[0].copy()
    du[2] = np.nan
    _, _, chol = ss.decompose_gram(du, arrs[1], arrs[2])
    bad = torch.isnan(chol).flatten(1).any(1).cpu().numpy()
    assert bad.tolist() == [False, False, True, False, False]


def test_state_space_full_size_properties_and_bandwidth():
    """65,536 chains x dim_y 1,000 x dim_u 3 (a stochastic-volatility-sized system):
    G (G^{-1} w) = w with G applied through lmult / rmult_by_jacob_constr, and
    (J J^T)^{-1} from the non-symmetric routine agrees with the symmetric one."""
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    ss = mlb.state_space
    n, dim_y, dim_u = 65536, 1000, 3
    g = torch.Generator(device="cuda").manual_seed(0)

    def rnd(rows):
        return _lib.soa_empty(n, rows).normal_(generator=g)

    J = ss.JacobBlocks.__new__(ss.JacobBlocks)
    J.n, J.dim_y, J.dim_u = n, dim_y, dim_u
    J.dc_du = rnd(dim_y * dim_u)
    J.dc_dv = rnd(dim_y).abs_().add_(0.5)
    J.dc_dn0 = rnd(dim_y).abs_().add_(0.5)
    J.dc_dn1 = rnd(dim_y - 1).abs_().mul_(0.4).add_(0.1)
    J.dx_dy = rnd(dim_y).abs_().add_(0.3)
    w = rnd(dim_y)
    a, b, chol = ss.decompose_gram(J)
    x = ss.lmult_by_inv_gram(J, None, None, None, a, b, chol, w)
    gx = ss.lmult_by_jacob_constr(J, None, None, None,
                                  ss.rmult_by_jacob_constr(J, None, None, None, x))
    assert float((gx - w).abs().max()) < 1e-9
    x2 = ss.lmult_by_inv_jacob_product(J, None, None, None, J, None, None, None, w)
    assert float((x2 - x).abs().max()) < 1e-9 * max(1.0, float(x.abs().max()))
    ld = ss.log_det_sqrt_gram(J, gram=(a, b, chol))
    assert bool(torch.isfinite(ld).all())

    def timed(fn, reps=5):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # algorithmic bytes: blocks read once, results written once
    cases = {
        "decompose_gram": (lambda: ss.decompose_gram(J), 8.0 * n * ((dim_u + 3) * dim_y + 2 * dim_y + dim_u**2)),
        "lmult_by_inv_gram": (lambda: ss.lmult_by_inv_gram(J, None, None, None, a, b, chol, w),
                              8.0 * n * ((dim_u + 4) * dim_y + dim_u**2)),
        "lmult_by_inv_jacob_product": (
            lambda: ss.lmult_by_inv_jacob_product(J, None, None, None, J, None, None, None, w),
            8.0 * n * ((2 * dim_u + 8) * dim_y)),
        "log_det_sqrt_gram": (lambda: ss.log_det_sqrt_gram(J, gram=(a, b, chol)), 8.0 * n * 3 * dim_y),
    }  # fmt: skip
    for name, (fn, alg_bytes) in cases.items():
        ms = timed(fn)
        print(f"state_space.{name}: {ms:.3f} ms per call, "
              f"{alg_bytes / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic")


def _sv_setup(dim_y, n, seed):
    rng = np.random.default_rng(seed)
    qs, y = oss.sv_on_manifold_states(rng, n, dim_y)
    return rng, qs, y


@pytest.mark.parametrize("dim_y", [1, 2, 25, 300])
def test_stochastic_volatility_constr_and_blocks_match_restatement(dim_y):
    import manifold_lifting_b200 as mlb

    n = 41
    rng, qs, y = _sv_setup(dim_y, n, dim_y)
    qs = qs + 0.01 * rng.normal(size=qs.shape)  # off the manifold: c != 0
    qs[:, 3 + dim_y :] = np.where(np.abs(qs[:, 3 + dim_y :]) < 0.05, 0.05, qs[:, 3 + dim_y :])
    qs[:, 3 + dim_y :] = np.abs(qs[:, 3 + dim_y :]) * np.sign(y)[None]  # y / n > 0
    model = mlb.state_space.StochasticVolatilityModel(y)
    ref = [oss.sv_jacob_constr_split_blocks(*oss.sv_split(qs[i], dim_y), y) for i in range(n)]
    c_ref = np.stack([r[1] for r in ref])
    _close(model.constr(qs), c_ref, 1e-13)
    J, c = model.jacob_constr_blocks(qs)
    _close(c, c_ref, 1e-13)
    _close(J.dc_du.reshape(n, dim_y, 3), np.stack([r[0][0] for r in ref]), 1e-13)
    _close(J.dc_dv, np.stack([r[0][1] for r in ref]), 1e-13)
    _close(J.dc_dn0, np.stack([r[0][2][0] for r in ref]), 1e-13)
    if dim_y > 1:
        _close(J.dc_dn1, np.stack([r[0][2][1] for r in ref]), 1e-13)
    _close(J.dx_dy, np.stack([r[0][3] for r in ref]), 1e-13)


@pytest.mark.parametrize("dim_y", [2, 40, 500])
def test_stochastic_volatility_projection_loops_match_restatement(dim_y):
    """The reference's quasi-Newton and Newton projection loops (systems.py:190-270) on the
    stochastic-volatility model: per chain the same iteration count, norms and projected
    position as the NumPy restatement; the projected points are on the manifold."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    n, dt = 24, 0.02
    rng, qs, y = _sv_setup(dim_y, n, 100 + dim_y)
    model = ss.StochasticVolatilityModel(y)
    J, c0 = model.jacob_constr_blocks(qs)
    assert float(c0.abs().max()) < 1e-12
    gram = ss.decompose_gram(J)
    # position step along a cotangent vector, as the integrator's h2_flow does
    p = rng.normal(size=qs.shape)
    Jp = ss.lmult_by_jacob_constr(J, None, None, None, p)
    p = torch.as_tensor(p).cuda() - ss.rmult_by_jacob_constr(
        J, None, None, None, ss.lmult_by_inv_gram(J, None, None, None, *gram, Jp))
    assert float(ss.lmult_by_jacob_constr(J, None, None, None, p).abs().max()) < 1e-8
    q1 = torch.as_tensor(qs).cuda() + dt * p
    q1_np = q1.cpu().numpy()
    jac_np = [oss.sv_jacob_constr_split_blocks(*oss.sv_split(qs[i], dim_y), y)[0] for i in range(n)]
    for solver in ("quasi_newton", "newton"):
        if solver == "quasi_newton":
            out = model.quasi_newton_projection(q1, J, gram, dt)
            ref = [oss.sv_quasi_newton_projection(q1_np[i], jac_np[i], oss.decompose_gram(*jac_np[i]),
                                                  dt, y) for i in range(n)]  # fmt: skip
        else:
            out = model.newton_projection(q1, J, dt)
            ref = [oss.sv_newton_projection(q1_np[i], jac_np[i], dt, y) for i in range(n)]
        q_new, mu, iters, norm_dq, err = out
        it_ref = np.array([r[2] for r in ref])
        # counts agree except where a norm sits within rounding of its tolerance
        same = iters.cpu().numpy() == it_ref
        assert same.mean() >= 0.9, (solver, iters.cpu().numpy(), it_ref)
        assert it_ref.max() < 50 and int(iters.max()) < 50
        qr = np.stack([r[0] for r in ref])
        mr = np.stack([r[1] for r in ref])
        assert np.abs(q_new.cpu().numpy()[same] - qr[same]).max() < 1e-10 * max(1.0, np.abs(qr).max())
        assert np.abs(mu.cpu().numpy()[same] - mr[same]).max() < 1e-8 * max(1.0, np.abs(mr).max())
        assert float(model.constr(q_new).abs().max()) < 1e-9
        assert float(err.max()) < 1e-9 and float(norm_dq.max()) < 1e-8


def test_stochastic_volatility_projection_flags_failures_per_chain():
    """A chain that cannot converge (max_iters too small / NaN position) stops by the
    reference's test without holding up or contaminating the others."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    dim_y, n, dt = 30, 8, 0.05
    rng, qs, y = _sv_setup(dim_y, n, 5)
    model = ss.StochasticVolatilityModel(y)
    J, _ = model.jacob_constr_blocks(qs)
    gram = ss.decompose_gram(J)
    q1 = qs + 1e-3 * rng.normal(size=qs.shape)
    q1[3, 5] = np.nan
    _, _, iters, norm_dq, err = model.quasi_newton_projection(q1, J, gram, dt, max_iters=50)
    assert int(iters[3]) == 1 and bool(torch.isnan(err[3]))
    ok = np.arange(n) != 3
    assert float(err.cpu().numpy()[ok].max()) < 1e-9
    _, _, iters2, _, _ = model.quasi_newton_projection(q1, J, gram, dt, max_iters=2)
    assert int(iters2.max()) <= 2 and int(iters2[0]) == 2


def test_stochastic_volatility_projection_full_size():
    """16,384 chains x dim_y 1,000: every chain lands on the manifold; time per iteration."""
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    ss = mlb.state_space
    n, dim_y, dt = 16384, 1000, 0.02
    rng = np.random.default_rng(11)
    u_ref = np.array([-0.5, np.log(0.3), 2.0])
    y = oss.sv_generate_y(u_ref, rng.normal(size=dim_y), oss.sv_observation_noise(rng, dim_y))
    model = ss.StochasticVolatilityModel(y)
    g = torch.Generator(device="cuda").manual_seed(1)
    q = _lib.soa_empty(n, model.dim_q).normal_(generator=g)
    q[:, 0] = -0.5 + 0.2 * q[:, 0]
    q[:, 1] = float(np.log(0.3)) + 0.2 * q[:, 1]
    q[:, 2] = 2.0 + 0.2 * q[:, 2]
    # lift onto the manifold: simulate x from (u, v), then n = y / exp(x / 2)
    sigma, phi = q[:, 1].exp(), -1.0 + 2.0 * torch.sigmoid(q[:, 2])
    x = q[:, 0] + sigma / (1 - phi**2).sqrt() * q[:, 3]
    yt = torch.as_tensor(y).cuda()
    for t in range(dim_y):
        if t > 0:
            x = q[:, 0] + phi * (x - q[:, 0]) + sigma * q[:, 3 + t]
        q[:, 3 + dim_y + t] = yt[t] / (x / 2).exp()
    J, c0 = model.jacob_constr_blocks(q)
    assert float(c0.abs().max()) < 1e-10
    gram = ss.decompose_gram(J)
    p = _lib.soa_empty(n, model.dim_q).normal_(generator=g)
    p = p - ss.rmult_by_jacob_constr(J, None, None, None, ss.lmult_by_inv_gram(
        J, None, None, None, *gram, ss.lmult_by_jacob_constr(J, None, None, None, p)))
    q1 = q + dt * p
    for name, run in (("quasi_newton", lambda: model.quasi_newton_projection(q1, J, gram, dt)),
                      ("newton", lambda: model.newton_projection(q1, J, dt))):
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        q_new, mu, iters, norm_dq, err = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        assert float(model.constr(q_new).abs().max()) < 1e-9
        assert float(err.max()) < 1e-9 and float(norm_dq.max()) < 1e-8
        print(f"sv {name}_projection: {n} chains x dim_y {dim_y}: {ms:.2f} ms, "
              f"{int(iters.max())} rounds (mean {float(iters.double().mean()):.2f} iterations per chain), "
              f"{n / (ms * 1e-3):.3g} projections/s")


@pytest.mark.parametrize("dim_y,dim_u", [(1, 1), (9, 3), (200, 8)])
def test_state_space_cotangent_projection(dim_y, dim_u):
    """p - J^T G^{-1} J p: against the restated closures, J p' = 0 afterwards, idempotent."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    rng = np.random.default_rng(50 + dim_y)
    n = 37
    chains, arrs = _batch(rng, n, dim_y, dim_u)
    J = ss.JacobBlocks(*arrs)
    gram = ss.decompose_gram(J)
    p = rng.normal(size=(n, dim_u + 2 * dim_y))
    out = ss.project_onto_cotangent_space(p, J, gram)
    ref = np.stack([
        p[i] - oss.rmult_by_jacob_constr(*chains[i], oss.lmult_by_inv_gram(
            *chains[i], *oss.decompose_gram(*chains[i]),
            oss.lmult_by_jacob_constr(*chains[i], p[i]))) for i in range(n)])  # fmt: skip
    _close(out, ref, 1e-12)
    assert float(ss.lmult_by_jacob_constr(J, None, None, None, out).abs().max()) < 1e-11
    _close(ss.project_onto_cotangent_space(out, J, gram), out.cpu().numpy(), 1e-12)


@pytest.mark.parametrize("solver", ["newton", "quasi_newton"])
def test_stochastic_volatility_constrained_position_step_matches_restatement(solver):
    """Position half of the constrained leapfrog step (projection, momentum update and
    re-projection, reversibility check) chain by chain against the restated integrator
    semantics; a chain pushed far off the manifold is flagged, not raised."""
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    ss = mlb.state_space
    dim_y, n, dt = 60, 20, 0.02
    rng, qs, y = _sv_setup(dim_y, n, 77)
    model = ss.StochasticVolatilityModel(y)
    J, _ = model.jacob_constr_blocks(qs)
    p = ss.project_onto_cotangent_space(rng.normal(size=qs.shape), J, ss.decompose_gram(J))
    p_np = p.cpu().numpy().copy()
    p_np[5] *= 400.0  # a step far too long for chain 5
    out = model.constrained_position_step(qs, p_np, dt, solver=solver)
    ref = [oss.sv_constrained_position_step(qs[i], p_np[i], dt, y, solver=solver) for i in range(n)]
    st_ref = np.array([r[2] for r in ref])
    st = out["status"].cpu().numpy()
    assert (st == st_ref).all(), (st, st_ref)
    assert st_ref[5] != _lib.ST_OK and (np.delete(st_ref, 5) == _lib.ST_OK).all()
    ok = st_ref == _lib.ST_OK
    ok &= out["iters_fwd"].cpu().numpy() == np.array([r[3] for r in ref])
    ok &= out["iters_rev"].cpu().numpy() == np.array([r[4] for r in ref])
    assert ok.sum() >= n - 3  # iteration counts may differ where a norm sits on a tolerance
    qr, pr = np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref])
    assert np.abs(out["pos"].cpu().numpy()[ok] - qr[ok]).max() < 1e-10 * max(1.0, np.abs(qr[ok]).max())
    assert np.abs(out["mom"].cpu().numpy()[ok] - pr[ok]).max() < 1e-8 * max(1.0, np.abs(pr[ok]).max())
    good = torch.as_tensor(st_ref == _lib.ST_OK).cuda()
    assert float(model.constr(out["pos"])[good].abs().max()) < 1e-9
    Jp = ss.lmult_by_jacob_constr(out["jacob_constr_blocks"], None, None, None, out["mom"])
    assert float(Jp[good].abs().max()) < 1e-9
    assert float(out["reverse_error"][good].max()) <= 2e-8


def _guarded(n, rows, g=3):
    """[n, rows] component-major view inside a larger array whose g rows before and after
    are NaN (compute-sanitizer is not available on the GPU pool: out-of-range reads that
    reach a result and out-of-range writes are caught with these canaries instead)."""
    from manifold_lifting_b200 import _lib

    big = _lib.soa_empty(n, rows + 2 * g)
    big.fill_(float("nan"))
    return big, big[:, g : g + rows]


@pytest.mark.parametrize("dim_y,dim_u", [(1, 1), (2, 2), (7, 3), (8, 3), (13, 4), (24, 5), (50, 8)])
def test_state_space_kernels_stay_inside_their_arrays(dim_y, dim_u):
    """Every input lives between NaN guard rows, every output is written between guard
    rows through the C ABI directly: results equal those from plain arrays bit for bit and
    the guards stay NaN.  dim_y around multiples of the prefetch depths (4, 6, 8)."""
    import ctypes as C

    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib
    from manifold_lifting_b200._lib import check, lib, stream_ptr

    ss = mlb.state_space
    rng = np.random.default_rng(900 + dim_y)
    n, g = 37, 3
    _, arrs = _batch(rng, n, dim_y, dim_u)
    _, arrs_r = _batch(rng, n, dim_y, dim_u)
    plain, plain_r = ss.JacobBlocks(*arrs), ss.JacobBlocks(*arrs_r)

    def guarded_blocks(src):
        J = ss.JacobBlocks.__new__(ss.JacobBlocks)
        J.n, J.dim_y, J.dim_u = n, dim_y, dim_u
        keep = []
        for name, rows in (("dc_du", dim_y * dim_u), ("dc_dv", dim_y), ("dc_dn0", dim_y),
                           ("dc_dn1", dim_y - 1), ("dx_dy", dim_y)):  # fmt: skip
            if rows == 0:
                setattr(J, name, None)
                continue
            big, view = _guarded(n, rows, g)
            view.copy_(getattr(src, name))
            keep.append(big)
            setattr(J, name, view)
        return J, keep

    J, keep = guarded_blocks(plain)
    Jr, keep_r = guarded_blocks(plain_r)
    vy = torch.as_tensor(rng.normal(size=(n, dim_y)))
    vq = torch.as_tensor(rng.normal(size=(n, dim_u + 2 * dim_y)))
    big_vy, gvy = _guarded(n, dim_y, g)
    gvy.copy_(vy)
    big_vq, gvq = _guarded(n, dim_u + 2 * dim_y, g)
    gvq.copy_(vq)

    def same(x, y):
        assert torch.equal(x, y), float((x - y).abs().max())

    a, b, chol = ss.decompose_gram(plain)
    a2, b2, chol2 = ss.decompose_gram(J)
    same(a, a2), same(b, b2), same(chol, chol2)
    same(ss.lmult_by_jacob_constr(plain, None, None, None, vq),
         ss.lmult_by_jacob_constr(J, None, None, None, gvq))
    same(ss.rmult_by_jacob_constr(plain, None, None, None, vy),
         ss.rmult_by_jacob_constr(J, None, None, None, gvy))
    same(ss.log_det_sqrt_gram(plain, gram=(a, b, chol)), ss.log_det_sqrt_gram(J, gram=(a, b, chol)))
    want_g = ss.lmult_by_inv_gram(plain, None, None, None, a, b, chol, vy)
    want_p = ss.lmult_by_inv_jacob_product(plain, None, None, None, plain_r, None, None, None, vy)
    want_c = ss.project_onto_cotangent_space(vq, plain, (a, b, chol))
    # guarded gram inputs and guarded outputs, through the C ABI
    big_a, ga = _guarded(n, max(dim_y - 1, 1), g)
    big_b, gb = _guarded(n, dim_y, g)
    big_c, gc = _guarded(n, dim_u * dim_u, g)
    if dim_y > 1:
        ga.copy_(a)
    gb.copy_(b), gc.copy_(chol.reshape(n, dim_u * dim_u))
    dims = (C.c_int64(n), C.c_int32(dim_y), C.c_int32(dim_u), C.c_int64(n))
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    cs, cr = J.c_struct(), Jr.c_struct()
    big_o, go = _guarded(n, dim_y, g)
    check(lib().mlb_ssm_lmult_by_inv_gram(*dims, C.byref(cs), ptr(ga) if dim_y > 1 else None,
                                          ptr(gb), ptr(gc), ptr(gvy), ptr(go), stream_ptr()))
    same(go, want_g)
    assert bool(torch.isnan(big_o[:, :g]).all()) and bool(torch.isnan(big_o[:, g + dim_y :]).all())
    big_o2, go2 = _guarded(n, dim_y, g)
    check(lib().mlb_ssm_lmult_by_inv_jacob_product(*dims, C.byref(cs), C.byref(cr), ptr(gvy),
                                                   ptr(go2), stream_ptr()))
    same(go2, want_p)
    assert bool(torch.isnan(big_o2[:, :g]).all()) and bool(torch.isnan(big_o2[:, g + dim_y :]).all())
    check(lib().mlb_ssm_project_onto_cotangent_space(
        *dims, C.byref(cs), ptr(ga) if dim_y > 1 else None, ptr(gb), ptr(gc), ptr(gvq),
        stream_ptr()))
    same(gvq, want_c)
    q_rows = dim_u + 2 * dim_y
    assert bool(torch.isnan(big_vq[:, :g]).all()) and bool(torch.isnan(big_vq[:, g + q_rows :]).all())
    # the inputs' guards are untouched as well
    for big, rows in ((big_vy, dim_y), (big_a, max(dim_y - 1, 1)), (big_b, dim_y)):
        assert bool(torch.isnan(big[:, :g]).all()) and bool(torch.isnan(big[:, g + rows :]).all())
    # tridiagonal solve (one right-hand side: the fused kernel) with guarded arrays
    big_x, gx = _guarded(n, dim_y, g)
    lower = ga if dim_y > 1 else None
    check(lib().mlb_tridiagonal_solve(C.c_int64(n), C.c_int32(dim_y), C.c_int32(1), C.c_int64(n),
                                      ptr(lower) if dim_y > 1 else None, ptr(gb),
                                      ptr(lower) if dim_y > 1 else None, ptr(gvy), ptr(gx),
                                      stream_ptr()))
    same(gx, mlb.linalg.tridiagonal_solve(a, b, a, vy) if dim_y > 1
         else mlb.linalg.tridiagonal_solve(None, b, None, vy))
    assert bool(torch.isnan(big_x[:, :g]).all()) and bool(torch.isnan(big_x[:, g + dim_y :]).all())


def _ref_sv():
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_sv.npz")
    return np.load(path)


@pytest.mark.parametrize("dim_y,dim_u", [(1, 1), (2, 3), (9, 2), (40, 3), (61, 5), (200, 8)])
def test_band_of_inverse_gram_matches_dense_inverse(dim_y, dim_u):
    """G^-1 dc_du, diag(G^-1) and the first off-diagonal of G^-1 from the two-sweep kernel
    against the dense inverse of the assembled J J^T (random blocks)."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    rng = np.random.default_rng(dim_y * 31 + dim_u)
    n = 37
    blocks = [oss.random_blocks(rng, dim_y, dim_u) for _ in range(n)]
    du = np.stack([b[0] for b in blocks])
    dv = np.stack([b[1] for b in blocks])
    dn0 = np.stack([b[2][0] for b in blocks])
    dn1 = np.stack([b[2][1] for b in blocks]).reshape(n, max(dim_y - 1, 0))
    dxdy = np.stack([b[3] for b in blocks])
    J = ss.JacobBlocks(du, dv, (dn0, dn1), dxdy)
    ru, gd, go = ss.band_of_inv_gram(J, ss.decompose_gram(J))
    for i in range(n):
        jac = oss.dense_jacobian(*blocks[i])
        ginv = np.linalg.inv(jac @ jac.T)
        scale = np.abs(ginv).max()
        assert np.abs(ru[i].cpu().numpy().reshape(dim_y, dim_u) - ginv @ blocks[i][0]).max() < 1e-11 * max(
            1.0, np.abs(ginv @ blocks[i][0]).max())
        assert np.abs(gd[i].cpu().numpy() - np.diag(ginv)).max() < 1e-11 * scale
        if dim_y > 1:
            assert np.abs(go[i].cpu().numpy() - np.diag(ginv, 1)).max() < 1e-11 * scale


def test_stochastic_volatility_log_det_gradient_matches_reference_source():
    """d log_det_sqrt_gram / dq of the stochastic-volatility system: the CUDA path against
    the outputs of the reference's own source (reverse-mode autodiff of systems.py:1233-1243
    on the torch-backed jax stand-in; tests/golden/ref_sv.npz) to the north-star's 1e-10,
    against the closed-form restatement on other sizes, and against central differences of
    the CUDA value."""
    import manifold_lifting_b200 as mlb

    ss = mlb.state_space
    g = _ref_sv()
    model = ss.StochasticVolatilityModel(g["y_obs"])
    for tag in ("closed", "auto"):
        (val, _), grad = model.val_and_grad_log_det_sqrt_gram(g[f"{tag}_q"])
        ref = g[f"{tag}_grad_logdet"]
        err = np.abs(grad.cpu().numpy() - ref).max(1) / np.abs(ref).max(1)
        assert err.max() < 1e-10, err
        assert np.abs(val.cpu().numpy() - g[f"{tag}_logdet"]).max() < 1e-10 * np.abs(g[f"{tag}_logdet"]).max()
    for dim_y in (1, 2, 7, 150):
        rng, qs, y = _sv_setup(dim_y, 12, 100 + dim_y)
        model = ss.StochasticVolatilityModel(y)
        (val, aux), grad = model.val_and_grad_log_det_sqrt_gram(qs)
        want = np.stack([oss.sv_grad_log_det_sqrt_gram(qs[i], y) for i in range(len(qs))])
        err = np.abs(grad.cpu().numpy() - want).max(1) / np.maximum(np.abs(want).max(1), 1.0)
        assert err.max() < 1e-9, (dim_y, err)
        # central differences of the CUDA value along random directions
        d = rng.normal(size=qs.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        eps = 1e-6
        fd = (model.log_det_sqrt_gram(qs + eps * d)[0] - model.log_det_sqrt_gram(qs - eps * d)[0]) / (2 * eps)
        an = (grad.cpu().numpy() * d).sum(1)
        assert np.abs(fd.cpu().numpy() - an).max() < 1e-6 * max(1.0, np.abs(an).max())


@pytest.mark.parametrize("solver", ["newton", "quasi_newton"])
def test_stochastic_volatility_leapfrog_step_matches_restatement(solver):
    """The whole constrained leapfrog step A(dt/2) B(dt) A(dt/2) for the state-space family
    chain by chain against the restated integrator: positions, momenta, statuses, iteration
    counts, Hamiltonians; the energy error is O(dt^2) and the step is time-reversible."""
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    ss = mlb.state_space
    dim_y, n, dt = 50, 16, 0.02
    rng, qs, y = _sv_setup(dim_y, n, 5)
    model = ss.StochasticVolatilityModel(y)
    J, _ = model.jacob_constr_blocks(qs)
    p = ss.project_onto_cotangent_space(rng.normal(size=qs.shape), J, ss.decompose_gram(J))
    p_np = p.cpu().numpy().copy()
    out = model.leapfrog_step(qs, p_np, dt, solver=solver)
    ref = [oss.sv_leapfrog_step(qs[i], p_np[i], dt, y, solver=solver) for i in range(n)]
    st_ref = np.array([r[2] for r in ref])
    assert (out["status"].cpu().numpy() == st_ref).all() and (st_ref == _lib.ST_OK).all()
    qr, pr = np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref])
    same = (out["iters_fwd"].cpu().numpy() == np.array([r[3] for r in ref])) & (
        out["iters_rev"].cpu().numpy() == np.array([r[4] for r in ref]))
    assert same.sum() >= n - 2
    assert np.abs(out["pos"].cpu().numpy()[same] - qr[same]).max() < 1e-10 * max(1.0, np.abs(qr).max())
    assert np.abs(out["mom"].cpu().numpy()[same] - pr[same]).max() < 1e-8 * max(1.0, np.abs(pr).max())
    h0 = np.array([r[5] for r in ref])
    h1 = np.array([r[6] for r in ref])
    assert np.abs(out["h_init"].cpu().numpy() - h0).max() < 1e-10 * np.abs(h0).max()
    assert np.abs(out["h_final"].cpu().numpy()[same] - h1[same]).max() < 1e-9 * np.abs(h1).max()
    # energy error O(dt^2): halving the step divides it by ~4 (loosely: by more than 2.5)
    e_full = (out["h_final"] - out["h_init"]).abs().cpu().numpy()
    half = model.leapfrog_step(qs, p_np, dt / 2, solver=solver)
    e_half = (half["h_final"] - half["h_init"]).abs().cpu().numpy()
    assert np.median(e_full / np.maximum(e_half, 1e-300)) > 2.5
    # reversibility of the whole step: flip the momentum, step, flip again
    back = model.leapfrog_step(out["pos"], -out["mom"], dt, solver=solver)
    assert float((back["pos"] - torch.as_tensor(qs).cuda()).abs().max()) < 1e-7
    assert float((-back["mom"] - torch.as_tensor(p_np).cuda()).abs().max()) < 1e-6
    good = torch.ones(n, dtype=torch.bool, device="cuda")
    assert float(model.constr(out["pos"])[good].abs().max()) < 1e-9
