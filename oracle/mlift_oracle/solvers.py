"""Oracle restatement of the mlift projection-solver wrappers (reference mlift/solvers.py).

Host glue around the three array-level loops: fetch the cached Jacobian blocks / Gram
components of `state_prev`, run the loop, add the iteration counts to the call
counters, and either write back `pos` / `mom` or raise `ConvergenceError`
(solvers.py:16-95, 98-169, 172-256).  TEST INFRASTRUCTURE ONLY.
"""
import math

import numpy as np
import torch


class IntegratorError(RuntimeError):
    """Stand-in for mici.errors.IntegratorError."""


class ConvergenceError(IntegratorError):
    """Stand-in for mici.errors.ConvergenceError."""


class NonReversibleStepError(IntegratorError):
    """Stand-in for mici.errors.NonReversibleStepError."""


def euclidean_norm(vct):
    # solvers.py:6-8
    return (vct**2).sum() ** 0.5


def maximum_norm(vct):
    # solvers.py:11-13
    return abs(vct).max()


def _bump(state, names, counts):
    if state.call_counts is None:
        return
    for name, k in zip(names, counts):
        state.call_counts[name] = state.call_counts.get(name, 0) + k


def _finish(state, q_new, mu, num_iters, norm_delta_q, error, c_tol, p_tol, div_tol, tag):
    # solvers.py:81-95
    if error < c_tol and norm_delta_q < p_tol:
        state.pos = np.array(q_new)
        if state.mom is not None:
            state.mom = state.mom - np.asarray(mu)
        return state
    if error > div_tol or math.isnan(error):
        raise ConvergenceError(
            f"{tag} iteration diverged on iteration {num_iters}. "
            f"Last |c|={error:.1e}, |dq|={norm_delta_q}."
        )
    raise ConvergenceError(
        f"{tag} iteration did not converge in {num_iters} iterations. "
        f"Last |c|={error:.1e}, |dq|={norm_delta_q}."
    )


def solve_projection_onto_manifold_quasi_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm,
):  # fmt: skip
    jac_prev = system.jacob_constr_blocks(state_prev)
    gram_prev = system.gram_components(state_prev)
    q_new, mu, n_it, ndq, err = system._quasi_newton_projection(
        state.pos, jac_prev, gram_prev, dt, constraint_tol, position_tol,
        divergence_tol, max_iters, norm,
    )  # fmt: skip
    _bump(state, ("constr", "lmult_by_pinv_jacob_constr"), (n_it, n_it))
    state.last_num_iters = n_it
    return _finish(state, q_new, mu, n_it, ndq, err, constraint_tol, position_tol,
                   divergence_tol, "Quasi-Newton")  # fmt: skip


def solve_projection_onto_manifold_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm,
):  # fmt: skip
    jac_prev = system.jacob_constr_blocks(state_prev)
    q_new, mu, n_it, ndq, err = system._newton_projection(
        state.pos, jac_prev, dt, constraint_tol, position_tol, divergence_tol,
        max_iters, norm,
    )  # fmt: skip
    _bump(
        state,
        ("constr", "jacob_constr_blocks", "rmult_by_jacob_constr",
         "lmult_by_inv_jacob_product"),
        (n_it,) * 4,
    )  # fmt: skip
    state.last_num_iters = n_it
    return _finish(state, q_new, mu, n_it, ndq, err, constraint_tol, position_tol,
                   divergence_tol, "Newton")  # fmt: skip


def solve_projection_onto_manifold_newton_with_line_search(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, max_line_search_iters=10, norm=maximum_norm,
):  # fmt: skip
    jac_prev = system.jacob_constr_blocks(state_prev)
    q_new, mu, n_it, ndq, err, n_c = system._newton_projection_with_line_search(
        state.pos, jac_prev, dt, constraint_tol, position_tol, divergence_tol,
        max_iters, max_line_search_iters, norm,
    )  # fmt: skip
    _bump(
        state,
        ("constr", "jacob_constr_blocks", "rmult_by_jacob_constr",
         "lmult_by_inv_jacob_product"),
        (n_c, n_it, n_it, n_it),
    )  # fmt: skip
    state.last_num_iters = n_it
    return _finish(state, q_new, mu, n_it, ndq, err, constraint_tol, position_tol,
                   divergence_tol, "Newton")  # fmt: skip


# names as exported by the reference package (mlift/__init__.py:1-3)
jitted_solve_projection_onto_manifold_quasi_newton = (
    solve_projection_onto_manifold_quasi_newton
)
jitted_solve_projection_onto_manifold_newton = solve_projection_onto_manifold_newton
jitted_solve_projection_onto_manifold_newton_with_line_search = (
    solve_projection_onto_manifold_newton_with_line_search
)

__all__ = [n for n in dir() if not n.startswith("_") and n not in ("math", "np", "torch")]
