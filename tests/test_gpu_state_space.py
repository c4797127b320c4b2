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
