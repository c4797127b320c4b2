"""Batched tridiagonal primitives of mlift.linalg on the GPU (csrc/mlb_linalg.cu, through
the C ABI) against the NumPy restatement of the reference's scans and against dense linear
algebra; full-size residual property and the achieved HBM rate."""
import numpy as np
import pytest
import torch

from mlift_oracle import linalg as ol

pytestmark = pytest.mark.gpu


def _systems(rng, n, dim, symmetric=False):
    a = rng.normal(size=(n, dim - 1))
    c = a.copy() if symmetric else rng.normal(size=(n, dim - 1))
    b = 2.5 + np.abs(rng.normal(size=(n, dim)))
    b[:, 1:] += np.abs(a)
    b[:, :-1] += np.abs(c)
    return a, b, c


@pytest.mark.parametrize("dim", [1, 2, 3, 17, 200, 1000])
def test_tridiagonal_solve_matches_restatement(dim):
    import manifold_lifting_b200 as mlb

    rng = np.random.default_rng(dim)
    n = 257
    a, b, c = _systems(rng, n, dim)
    d = rng.normal(size=(n, dim))
    x = mlb.linalg.tridiagonal_solve(a, b, c, d).cpu().numpy()
    ref = np.stack([ol.tridiagonal_solve(a[i], b[i], c[i], d[i]) for i in range(n)])
    assert np.abs(x - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())
    # several right-hand sides with one factorisation (the reference vmaps over them)
    d3 = rng.normal(size=(n, dim, 3))
    x3 = mlb.linalg.tridiagonal_solve(a, b, c, d3).cpu().numpy()
    for r in range(3):
        ref = np.stack([ol.tridiagonal_solve(a[i], b[i], c[i], d3[i, :, r]) for i in range(0, n, 16)])
        assert np.abs(x3[::16, :, r] - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dim", [1, 2, 5, 300])
def test_tridiagonal_log_det_matches_restatement_and_slogdet(dim):
    import manifold_lifting_b200 as mlb

    rng = np.random.default_rng(100 + dim)
    n = 130
    a, b, _ = _systems(rng, n, dim, symmetric=True)
    out = mlb.linalg.tridiagonal_pos_def_log_det(a, b).cpu().numpy()
    ref = np.array([ol.tridiagonal_pos_def_log_det(a[i], b[i]) for i in range(n)])
    assert np.abs(out - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    dense = np.array([np.linalg.slogdet(np.diag(b[i]) + np.diag(a[i], -1) + np.diag(a[i], 1))[1]
                      for i in range(0, n, 13)])  # fmt: skip
    assert np.abs(out[::13] - dense).max() < 1e-10 * max(1.0, np.abs(dense).max())


def test_tridiagonal_full_size_residual_and_bandwidth():
    """65,536 chains x dim 1,000 (the state-space models' dim_y): T x = d to rounding, and
    the two kernels together move their algorithmic bytes at a sizeable share of HBM."""
    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    n, dim = 65536, 1000
    g = torch.Generator(device="cuda").manual_seed(0)
    a = _lib.soa_empty(n, dim - 1, "cuda").normal_(generator=g)
    c = _lib.soa_empty(n, dim - 1, "cuda").normal_(generator=g)
    b = _lib.soa_empty(n, dim, "cuda").normal_(generator=g).abs_().add_(2.5)
    b[:, 1:] += a.abs()
    b[:, :-1] += c.abs()
    d = _lib.soa_empty(n, dim, "cuda").normal_(generator=g)
    x = mlb.linalg.tridiagonal_solve(a, b, c, d)
    r = b * x - d
    r[:, 1:] += a * x[:, :-1]
    r[:, :-1] += c * x[:, 1:]
    assert float(r.abs().max()) < 1e-12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mlb.linalg.tridiagonal_solve(a, b, c, d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    alg_bytes = 8.0 * n * (5 * dim - 2)  # read a, b, c, d; write x
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    print(f"tridiagonal_solve: {ms:.3f} ms per call, {gbs:.0f} GB/s algorithmic")
    assert gbs > 500.0
