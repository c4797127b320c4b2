"""Projection solvers - batched CUDA drop-in for `mlift.solvers` (reference
mlift/solvers.py).

Same function names, keyword arguments and defaults as the reference
(solvers.py:16-26, 98-108, 172-183).  Each call projects ALL chains of `state` in one
kernel launch.  Per-chain failures cannot be exceptions in a batched kernel: the solver
writes `state.status` (int32 per chain, 0 = converged) and leaves failed chains'
`pos` / `mom` untouched, exactly as if the reference had raised for that chain.  With a
single chain, or `raise_on_failure=True`, the reference's `ConvergenceError` is raised
(diverged vs not-converged messages as solvers.py:86-95).
"""
import torch

from . import _lib
from .errors import ConvergenceError


def euclidean_norm(vct):
    """Euclidean (L-2) norm over the last axis (solvers.py:6-8)."""
    return (vct**2).sum(-1) ** 0.5


def maximum_norm(vct):
    """Maximum (L-infinity) norm over the last axis (solvers.py:11-13)."""
    return abs(vct).max(-1).values if isinstance(vct, torch.Tensor) else abs(vct).max(-1)


def _bump(state, names, counts):
    # solvers.py:74-80: iteration counts are added to the per-state call counters
    if state._call_counts is None:
        return
    for name, k in zip(names, counts):
        state._call_counts[name] = state._call_counts.get(name, 0) + int(k)


def _finish(state, q_new, mu, num_iters, norm_delta_q, error, status, tag, raise_on_failure):
    ok = status == _lib.ST_OK
    state.status, state.last_num_iters = status, num_iters
    state.last_norm_delta_q, state.last_error = norm_delta_q, error
    all_ok = bool(ok.all())
    if not all_ok and (raise_on_failure or state.n_chain == 1):
        bad = int(torch.nonzero(~ok)[0])
        e, d, it = float(error[bad]), float(norm_delta_q[bad]), int(num_iters[bad])
        if int(status[bad]) == _lib.ST_DIVERGED:
            raise ConvergenceError(f"{tag} iteration diverged on iteration {it}. "
                                   f"Last |c|={e:.1e}, |δq|={d}.")  # fmt: skip
        raise ConvergenceError(f"{tag} iteration did not converge in {it} iterations. "
                               f"Last |c|={e:.1e}, |δq|={d}.")  # fmt: skip
    # solvers.py:81-85: pos <- q_, mom -= mu  (only for converged chains)
    if all_ok:
        new_pos = q_new
        if state.mom is not None:
            state.mom -= mu
    else:
        new_pos = _lib.soa_clone(state.pos)
        new_pos[ok] = q_new[ok]
        if state.mom is not None:
            state.mom[ok] -= mu[ok]
    state.pos = new_pos
    state.status = status
    return state


def jitted_solve_projection_onto_manifold_quasi_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm, raise_on_failure=False,
):  # fmt: skip
    """Symmetric quasi-Newton projection (solvers.py:16-95): fixed Gram of `state_prev`."""
    jac_prev = system.jacob_constr_blocks(state_prev)
    gram_prev = system.gram_components(state_prev)
    q_, mu, n_it, ndq, err, _, status = system._solve(
        _lib.SOLVER_QUASI_NEWTON, state.pos, jac_prev, gram_prev, dt, constraint_tol,
        position_tol, divergence_tol, max_iters, norm)  # fmt: skip
    tot = n_it.sum()
    _bump(state, ("constr", "lmult_by_pinv_jacob_constr"), (tot, tot))
    return _finish(state, q_, mu, n_it, ndq, err, status, "Quasi-Newton", raise_on_failure)


def jitted_solve_projection_onto_manifold_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm, raise_on_failure=False,
):  # fmt: skip
    """Full Newton projection (solvers.py:98-169): Jacobian re-evaluated every iteration."""
    jac_prev = system.jacob_constr_blocks(state_prev)
    q_, mu, n_it, ndq, err, _, status = system._solve(
        _lib.SOLVER_NEWTON, state.pos, jac_prev, None, dt, constraint_tol, position_tol,
        divergence_tol, max_iters, norm)  # fmt: skip
    tot = n_it.sum()
    _bump(state, ("constr", "jacob_constr_blocks", "rmult_by_jacob_constr",
                  "lmult_by_inv_jacob_product"), (tot,) * 4)  # fmt: skip
    return _finish(state, q_, mu, n_it, ndq, err, status, "Newton", raise_on_failure)


def jitted_solve_projection_onto_manifold_newton_with_line_search(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, max_line_search_iters=10, norm=maximum_norm,
    raise_on_failure=False,
):  # fmt: skip
    """Newton projection with backtracking line search (solvers.py:172-256)."""
    jac_prev = system.jacob_constr_blocks(state_prev)
    q_, mu, n_it, ndq, err, n_c, status = system._solve(
        _lib.SOLVER_NEWTON_LS, state.pos, jac_prev, None, dt, constraint_tol, position_tol,
        divergence_tol, max_iters, norm, max_line_search_iters)  # fmt: skip
    tot = n_it.sum()
    _bump(state, ("constr", "jacob_constr_blocks", "rmult_by_jacob_constr",
                  "lmult_by_inv_jacob_product"), (n_c.sum(), tot, tot, tot))  # fmt: skip
    return _finish(state, q_, mu, n_it, ndq, err, status, "Newton", raise_on_failure)


def solve_projection_onto_manifold_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm, raise_on_failure=False,
):  # fmt: skip
    """Mici's own Newton solver semantics (convergence tested before the update), as
    used by the reference's toy-model script (toy_2d_model.py:170-174)."""
    jac_prev = system.jacob_constr_blocks(state_prev)
    q_, mu, n_it, ndq, err, _, status = system._solve(
        _lib.SOLVER_MICI_NEWTON, state.pos, jac_prev, None, dt, constraint_tol,
        position_tol, divergence_tol, max_iters, norm)  # fmt: skip
    return _finish(state, q_, mu, n_it, ndq, err, status, "Newton", raise_on_failure)


def solve_projection_onto_manifold_quasi_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm, raise_on_failure=False,
):  # fmt: skip
    """Mici's symmetric quasi-Newton solver semantics (toy_2d_model.py:170-174)."""
    jac_prev = system.jacob_constr_blocks(state_prev)
    gram_prev = system.gram_components(state_prev)
    q_, mu, n_it, ndq, err, _, status = system._solve(
        _lib.SOLVER_MICI_QUASI_NEWTON, state.pos, jac_prev, gram_prev, dt, constraint_tol,
        position_tol, divergence_tol, max_iters, norm)  # fmt: skip
    return _finish(state, q_, mu, n_it, ndq, err, status, "Quasi-Newton", raise_on_failure)


SOLVER_IDS = {
    jitted_solve_projection_onto_manifold_quasi_newton: _lib.SOLVER_QUASI_NEWTON,
    jitted_solve_projection_onto_manifold_newton: _lib.SOLVER_NEWTON,
    jitted_solve_projection_onto_manifold_newton_with_line_search: _lib.SOLVER_NEWTON_LS,
    solve_projection_onto_manifold_newton: _lib.SOLVER_MICI_NEWTON,
    solve_projection_onto_manifold_quasi_newton: _lib.SOLVER_MICI_QUASI_NEWTON,
}

__all__ = [
    "euclidean_norm", "maximum_norm",
    "jitted_solve_projection_onto_manifold_quasi_newton",
    "jitted_solve_projection_onto_manifold_newton",
    "jitted_solve_projection_onto_manifold_newton_with_line_search",
]  # fmt: skip
