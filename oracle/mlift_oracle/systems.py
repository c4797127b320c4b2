"""Oracle restatement of the mlift constrained systems (reference mlift/systems.py).

Restates, single chain at a time exactly like the reference:
  * the matrix-free abstract system and its three projection loops
    (systems.py:148-347; loops :190-228 quasi-Newton, :230-270 Newton,
    :272-326 Newton + backtracking line search),
  * the Mici-facing cached methods (systems.py:369-471),
  * `IndependentAdditiveNoiseModelSystem` (systems.py:474-654) and
    `HierarchicalLatentVariableModelSystem` (systems.py:657-807), both the dense
    Cholesky and the Woodbury (capacitance-matrix) branches.

`torch.func.{jacfwd,jacrev,jvp,grad}` in fp64 stand in for the JAX transforms of the
same names; Python `while` loops stand in for `lax.while_loop` with the same carried
tuple, initial values and continuation predicate.  NumPy arrays cross the system
boundary as they do in the reference (systems.py:36-49).

TEST INFRASTRUCTURE ONLY - see package docstring.  PARITY UNPINNED (no reference
fixtures exist for this path).
"""
import math

import numpy as np
import torch
from torch.func import grad as _grad
from torch.func import jacfwd, jacrev, jvp

from .solvers import maximum_norm


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)


def _np(tree):
    """JAX->NumPy pytree conversion analogue (systems.py:36-49)."""
    if isinstance(tree, torch.Tensor):
        return tree.detach().numpy().copy()
    if isinstance(tree, (tuple, list)):
        return type(tree)(_np(s) for s in tree)
    return tree


def _tt(tree):
    if isinstance(tree, np.ndarray):
        return torch.as_tensor(tree, dtype=torch.float64)
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tt(s) for s in tree)
    return tree


def standard_normal_neg_log_dens(q):
    return 0.5 * (q**2).sum()


def cho_solve_lower(chol, b):
    return torch.cholesky_solve(b.reshape(-1, 1), chol, upper=False).reshape(-1)


class AbstractDifferentiableGenerativeModelSystem:
    """Matrix-free constrained system (systems.py:148-471).

    Subclasses provide the per-model closures as methods prefixed `f_` operating on
    torch tensors.  `neg_log_dens` is a torch function of q (the reference passes
    jitted value / gradient callables built by systems.py:52-87).
    """

    def __init__(self, neg_log_dens=standard_normal_neg_log_dens):
        self._nld_t = neg_log_dens

    # ---- default compositions (systems.py:173-188)
    def f_lmult_by_pinv_jacob_constr(self, jac, gram, vct):
        return self.f_rmult_by_jacob_constr(*jac, self.f_lmult_by_inv_gram(*jac, *gram, vct))

    def f_normal_space_component(self, jac, gram, vct):
        return self.f_lmult_by_pinv_jacob_constr(
            jac, gram, self.f_lmult_by_jacob_constr(*jac, vct)
        )

    # ---- shared loop predicate (systems.py:215-223, 257-265, 313-321)
    @staticmethod
    def _keep_iterating(i, norm_delta_q, error, c_tol, p_tol, div_tol, max_iters):
        diverged = (error > div_tol) or math.isnan(error)
        converged = (error < c_tol) and (norm_delta_q < p_tol)
        return not ((i >= max_iters) or diverged or converged)

    # ---- projection loops, array level ("jitted" attributes, systems.py:340-346)
    def _quasi_newton_projection(
        self, q, jac_prev, gram_prev, dt, constraint_tol, position_tol, divergence_tol,
        max_iters, norm=maximum_norm,
    ):  # fmt: skip
        q = _t(q).clone()
        jac_prev, gram_prev = _tt(jac_prev), _tt(gram_prev)
        mu, i, norm_dq, error = torch.zeros_like(q), 0, math.inf, -1.0
        while self._keep_iterating(
            i, norm_dq, error, constraint_tol, position_tol, divergence_tol, max_iters
        ):
            c = self.f_constr(q)
            error = float(norm(c))
            delta = self.f_lmult_by_pinv_jacob_constr(jac_prev, gram_prev, c)
            mu = mu + delta
            q = q - delta
            i += 1
            norm_dq = float(norm(delta))
        return _np(q), _np(mu / dt), i, norm_dq, error

    def _newton_projection(
        self, q, jac_prev, dt, constraint_tol, position_tol, divergence_tol, max_iters,
        norm=maximum_norm,
    ):  # fmt: skip
        q = _t(q).clone()
        jac_prev = _tt(jac_prev)
        mu, i, norm_dq, error = torch.zeros_like(q), 0, math.inf, -1.0
        while self._keep_iterating(
            i, norm_dq, error, constraint_tol, position_tol, divergence_tol, max_iters
        ):
            jac, c = self.f_jacob_constr_blocks(q)
            error = float(norm(c))
            delta = self.f_rmult_by_jacob_constr(
                *jac_prev, self.f_lmult_by_inv_jacob_product(*jac, *jac_prev, c)
            )
            mu = mu + delta
            q = q - delta
            i += 1
            norm_dq = float(norm(delta))
        return _np(q), _np(mu / dt), i, norm_dq, error

    def _newton_projection_with_line_search(
        self, q, jac_prev, dt, constraint_tol, position_tol, divergence_tol, max_iters,
        max_line_search_iters, norm=maximum_norm,
    ):  # fmt: skip
        q0 = _t(q).clone()
        jac_prev = _tt(jac_prev)
        q, norm_dq, error, i, n_constr = q0, math.inf, -1.0, 0, 0
        while self._keep_iterating(
            i, norm_dq, error, constraint_tol, position_tol, divergence_tol, max_iters
        ):
            jac, c = self.f_jacob_constr_blocks(q)
            err_here = float(norm(c))
            direction = -self.f_rmult_by_jacob_constr(
                *jac_prev, self.f_lmult_by_inv_jacob_product(*jac, *jac_prev, c)
            )
            # first trial unconditionally, then halve while the residual grew
            step, j = 1.0, 0
            new_q = q + step * direction
            new_err = float(norm(self.f_constr(new_q)))
            step, j = step * 0.5, j + 1
            while j < max_line_search_iters and new_err > err_here:
                new_q = q + step * direction
                new_err = float(norm(self.f_constr(new_q)))
                step, j = step * 0.5, j + 1
            norm_dq = float(norm(new_q - q))
            q, error, i, n_constr = new_q, new_err, i + 1, n_constr + j
        return _np(q), _np((q0 - q) / dt), i, norm_dq, error, n_constr

    # ---- array-level callables mirroring the jit table (systems.py:328-339)
    def _constr(self, q):
        return _np(self.f_constr(_t(q)))

    def _jacob_constr_blocks(self, q):
        return _np(self.f_jacob_constr_blocks(_t(q)))

    def _decompose_gram(self, *jac):
        return _np(self.f_decompose_gram(*_tt(jac)))

    def _log_det_sqrt_gram(self, q):
        return _np(self.f_log_det_sqrt_gram(_t(q)))

    def _val_and_grad_log_det_sqrt_gram(self, q):
        g, (val, aux) = _grad(
            lambda x: (lambda r: (r[0], r))(self.f_log_det_sqrt_gram(x)), has_aux=True
        )(_t(q))
        return _np(((val, aux), g))

    def _lmult_by_jacob_constr(self, *args):
        return _np(self.f_lmult_by_jacob_constr(*_tt(args)))

    def _rmult_by_jacob_constr(self, *args):
        return _np(self.f_rmult_by_jacob_constr(*_tt(args)))

    def _lmult_by_pinv_jacob_constr(self, jac, gram, vct):
        return _np(self.f_lmult_by_pinv_jacob_constr(_tt(jac), _tt(gram), _t(vct)))

    def _lmult_by_inv_jacob_product(self, *args):
        return _np(self.f_lmult_by_inv_jacob_product(*_tt(args)))

    def _normal_space_component(self, jac, gram, vct):
        return _np(self.f_normal_space_component(_tt(jac), _tt(gram), _t(vct)))

    # ---- Mici-facing, cached-in-state methods (systems.py:369-471)
    def neg_log_dens(self, state):
        if "nld" not in state.cache:
            g, v = _grad(lambda x: (lambda r: (r, r))(self._nld_t(x)), has_aux=True)(
                _t(state.pos)
            )
            state.cache["nld"], state.cache["grad_nld"] = float(v), _np(g)
        return state.cache["nld"]

    def grad_neg_log_dens(self, state):
        self.neg_log_dens(state)
        return state.cache["grad_nld"]

    def constr(self, state):
        if "constr" not in state.cache:
            state.cache["constr"] = self._constr(state.pos)
        return state.cache["constr"]

    def jacob_constr_blocks(self, state):
        if "jac" not in state.cache:
            jac, c = self._jacob_constr_blocks(state.pos)
            state.cache["jac"], state.cache["constr"] = jac, c
        return state.cache["jac"]

    def gram_components(self, state):
        if "gram" not in state.cache:
            state.cache["gram"] = self._decompose_gram(*self.jacob_constr_blocks(state))
        return state.cache["gram"]

    def log_det_sqrt_gram(self, state):
        if "logdet" not in state.cache:
            val, (c, jac, gram) = self._log_det_sqrt_gram(state.pos)
            state.cache.update(logdet=float(val), constr=c, jac=jac, gram=gram)
        return state.cache["logdet"]

    def grad_log_det_sqrt_gram(self, state):
        if "grad_logdet" not in state.cache:
            (val, (c, jac, gram)), g = self._val_and_grad_log_det_sqrt_gram(state.pos)
            state.cache.update(
                grad_logdet=g, logdet=float(val), constr=c, jac=jac, gram=gram
            )
        return state.cache["grad_logdet"]

    def h1(self, state):
        return self.neg_log_dens(state) + self.log_det_sqrt_gram(state)

    def dh1_dpos(self, state):
        return self.grad_neg_log_dens(state) + self.grad_log_det_sqrt_gram(state)

    def h2(self, state):
        return 0.5 * state.mom @ state.mom

    def h(self, state):
        return self.h1(state) + self.h2(state)

    def dh2_dmom(self, state):
        return state.mom

    def h1_flow(self, state, dt):
        # Mici System.h1_flow (not in the reference tree)
        state.mom = state.mom - dt * self.dh1_dpos(state)

    def h2_flow(self, state, dt):
        state.pos = state.pos + dt * self.dh2_dmom(state)

    def normal_space_component(self, state, vct):
        return self._normal_space_component(
            self.jacob_constr_blocks(state), self.gram_components(state), vct
        )

    def lmult_by_jacob_constr(self, state, vct):
        return self._lmult_by_jacob_constr(*self.jacob_constr_blocks(state), vct)

    def rmult_by_jacob_constr(self, state, vct):
        return self._rmult_by_jacob_constr(*self.jacob_constr_blocks(state), vct)

    def lmult_by_pinv_jacob_constr(self, state, vct):
        return self._lmult_by_pinv_jacob_constr(
            self.jacob_constr_blocks(state), self.gram_components(state), vct
        )

    def lmult_by_inv_jacob_product(self, state_1, state_2, vct):
        return self._lmult_by_inv_jacob_product(
            *self.jacob_constr_blocks(state_1), *self.jacob_constr_blocks(state_2), vct
        )

    def project_onto_cotangent_space(self, mom, state):
        return mom - self.normal_space_component(state, mom)

    def sample_momentum(self, state, rng):
        mom = rng.standard_normal(state.pos.shape)
        return self.project_onto_cotangent_space(mom, state)


class IndependentAdditiveNoiseModelSystem(AbstractDifferentiableGenerativeModelSystem):
    """y = F(u) + s(u) T(n); J = (dy_du dense, dy_dn diagonal) (systems.py:474-654)."""

    def __init__(
        self, generate_y, data, dim_u, neg_log_dens=standard_normal_neg_log_dens,
        force_full_gram_cholesky=False,
    ):  # fmt: skip
        super().__init__(neg_log_dens)
        self.generate_y, self.data, self.dim_u = generate_y, data, dim_u
        self.y_obs = _t(data["y_obs"])
        self.dim_y = int(self.y_obs.shape[0])
        self.dense = dim_u >= self.dim_y or force_full_gram_cholesky

    def _split(self, q):
        return q[: self.dim_u], q[self.dim_u :]

    def f_constr(self, q):
        u, n = self._split(q)
        return self.generate_y(u, n, self.data) - self.y_obs

    def f_jacob_constr_blocks(self, q):
        u, n = self._split(q)
        jac_fn = jacfwd if self.dim_u <= self.dim_y else jacrev
        dy_du = jac_fn(lambda u_: self.generate_y(u_, n, self.data))(u)
        y, dy_dn = jvp(
            lambda n_: self.generate_y(u, n_, self.data), (n,), (torch.ones(self.dim_y),)
        )
        return (dy_du, dy_dn), y - self.y_obs

    def f_lmult_by_jacob_constr(self, dy_du, dy_dn, vct):
        return dy_du @ vct[: self.dim_u] + dy_dn * vct[self.dim_u :]

    def f_rmult_by_jacob_constr(self, dy_du, dy_dn, vct):
        return torch.cat((vct @ dy_du, vct * dy_dn))

    def f_decompose_gram(self, dy_du, dy_dn):
        if self.dense:
            return (torch.linalg.cholesky(dy_du @ dy_du.T + torch.diag(dy_dn**2)),)
        d = dy_dn**2
        cap = torch.eye(self.dim_u) + dy_du.T @ (dy_du / d[:, None])
        return torch.linalg.cholesky(cap), d

    def f_lmult_by_inv_gram(self, dy_du, dy_dn, *gram_and_vct):
        if self.dense:
            chol, vct = gram_and_vct
            return cho_solve_lower(chol, vct)
        chol_cap, d, vct = gram_and_vct
        return (vct - dy_du @ cho_solve_lower(chol_cap, dy_du.T @ (vct / d))) / d

    def f_lmult_by_inv_jacob_product(self, dy_du_l, dy_dn_l, dy_du_r, dy_dn_r, vct):
        d = dy_dn_l * dy_dn_r
        if self.dense:
            return torch.linalg.solve(dy_du_l @ dy_du_r.T + torch.diag(d), vct)
        cap = torch.eye(self.dim_u) + dy_du_r.T @ (dy_du_l / d[:, None])
        return (vct - dy_du_l @ torch.linalg.solve(cap, dy_du_r.T @ (vct / d))) / d

    def f_log_det_sqrt_gram(self, q):
        jac, c = self.f_jacob_constr_blocks(q)
        gram = self.f_decompose_gram(*jac)
        val = torch.log(torch.diagonal(gram[0])).sum()
        if not self.dense:
            val = val + torch.log(gram[1]).sum() / 2
        return val, (c, jac, gram)


class HierarchicalLatentVariableModelSystem(AbstractDifferentiableGenerativeModelSystem):
    """y = g(u, v, n); J = (dy_du dense, dy_dv diag, dy_dn diag) (systems.py:657-807)."""

    def __init__(self, generate_y, data, dim_u, neg_log_dens=standard_normal_neg_log_dens):
        super().__init__(neg_log_dens)
        self.generate_y, self.data, self.dim_u = generate_y, data, dim_u
        self.y_obs = _t(data["y_obs"])
        self.dim_y = int(self.y_obs.shape[0])
        self.dense = dim_u >= self.dim_y

    def _split(self, q):
        U, Y = self.dim_u, self.dim_y
        return q[:U], q[U : U + Y], q[U + Y :]

    def f_constr(self, q):
        u, v, n = self._split(q)
        return self.generate_y(u, v, n, self.data) - self.y_obs

    def f_jacob_constr_blocks(self, q):
        u, v, n = self._split(q)
        one = torch.ones(self.dim_y)
        jac_fn = jacfwd if self.dim_u <= self.dim_y else jacrev
        dy_du = jac_fn(lambda u_: self.generate_y(u_, v, n, self.data))(u)
        _, dy_dv = jvp(lambda v_: self.generate_y(u, v_, n, self.data), (v,), (one,))
        y, dy_dn = jvp(lambda n_: self.generate_y(u, v, n_, self.data), (n,), (one,))
        return (dy_du, dy_dv, dy_dn), y - self.y_obs

    def f_lmult_by_jacob_constr(self, dy_du, dy_dv, dy_dn, vct):
        vu, vv, vn = self._split(vct)
        return dy_du @ vu + dy_dv * vv + dy_dn * vn

    def f_rmult_by_jacob_constr(self, dy_du, dy_dv, dy_dn, vct):
        return torch.cat((vct @ dy_du, vct * dy_dv, vct * dy_dn))

    def f_decompose_gram(self, dy_du, dy_dv, dy_dn):
        d = dy_dv**2 + dy_dn**2
        if self.dense:
            return (torch.linalg.cholesky(dy_du @ dy_du.T + torch.diag(d)),)
        cap = torch.eye(self.dim_u) + dy_du.T @ (dy_du / d[:, None])
        return torch.linalg.cholesky(cap), d

    def f_lmult_by_inv_gram(self, dy_du, dy_dv, dy_dn, *gram_and_vct):
        if self.dense:
            chol, vct = gram_and_vct
            return cho_solve_lower(chol, vct)
        chol_cap, d, vct = gram_and_vct
        return (vct - dy_du @ cho_solve_lower(chol_cap, dy_du.T @ (vct / d))) / d

    def f_lmult_by_inv_jacob_product(
        self, dy_du_l, dy_dv_l, dy_dn_l, dy_du_r, dy_dv_r, dy_dn_r, vct
    ):
        d = dy_dv_l * dy_dv_r + dy_dn_l * dy_dn_r
        if self.dense:
            return torch.linalg.solve(dy_du_l @ dy_du_r.T + torch.diag(d), vct)
        cap = torch.eye(self.dim_u) + dy_du_r.T @ (dy_du_l / d[:, None])
        return (vct - dy_du_l @ torch.linalg.solve(cap, dy_du_r.T @ (vct / d))) / d

    def f_log_det_sqrt_gram(self, q):
        jac, c = self.f_jacob_constr_blocks(q)
        gram = self.f_decompose_gram(*jac)
        val = torch.log(torch.diagonal(gram[0])).sum()
        if not self.dense:
            val = val + torch.log(gram[1]).sum() / 2
        return val, (c, jac, gram)
