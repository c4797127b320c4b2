"""Oracle restatement of the Mici 0.2.1 pieces the mlift hot path is driven by.

Mici's source is NOT under /root/reference (third-party dependency pinned at
`mici==0.2.1`, reference environment.yml:17).  What is restated here is Mici's published
behaviour as constrained by the reference's call sites:
  * integrator construction / attributes:  mlift/example_models/utils.py:317-327,
    mlift/adapters.py:45-47,56-72, toy_2d_model.py:183-185
  * solver signature and side effects:      mlift/solvers.py:16-26,81-95
  * system methods Mici calls:              mlift/systems.py:407-471
  * dense constrained system (toy model):   toy_2d_model.py:175-182
  * static Metropolis HMC accept rule:      notebook `simulate_proposal` (cell 44)
  * adapters:                               utils.py:281-285,328-334, adapters.py:5-72
Integrator semantics were RESTATED, NOT EXECUTED.
TEST INFRASTRUCTURE ONLY - see package docstring.
"""
import math

import numpy as np
import torch
from torch.func import grad as _grad
from torch.func import jacfwd

from .solvers import (
    ConvergenceError,
    IntegratorError,
    NonReversibleStepError,
    maximum_norm,
)


class ChainState:
    """pos / mom / dir with a per-position cache (Mici ChainState + cache_in_state)."""

    def __init__(self, pos, mom=None, dir=1, call_counts=None):
        self._pos = np.array(pos, dtype=np.float64)
        self.mom = None if mom is None else np.array(mom, dtype=np.float64)
        self.dir = dir
        self.cache = {}
        self.call_counts = call_counts
        self.last_num_iters = 0

    @property
    def pos(self):
        return self._pos

    @pos.setter
    def pos(self, value):
        self._pos = value
        self.cache = {}  # every cached quantity depends on pos only

    def copy(self):
        new = ChainState(self._pos, self.mom, self.dir, self.call_counts)
        new.cache = dict(self.cache)
        return new


class DenseConstrainedEuclideanMetricSystem:
    """Dense-Jacobian constrained system with identity metric (toy configs).

    h1 = nld + 0.5 log|J J^T| (dens_wrt_hausdorff=False), gradient of the log-det term
    from second derivatives of the constraint; cotangent projection p -= J^T G^-1 J p.
    """

    def __init__(self, neg_log_dens, constr):
        self._nld_t, self._constr_t = neg_log_dens, constr

    def _fill(self, state):
        if "jac" in state.cache:
            return
        q = torch.as_tensor(state.pos)
        state.cache["constr"] = self._constr_t(q).detach().numpy()
        state.cache["jac"] = jacfwd(self._constr_t)(q).detach().numpy()

    def constr(self, state):
        self._fill(state)
        return state.cache["constr"]

    def jacob_constr(self, state):
        self._fill(state)
        return state.cache["jac"]

    def gram(self, state):
        J = self.jacob_constr(state)
        return J @ J.T

    def neg_log_dens(self, state):
        return float(self._nld_t(torch.as_tensor(state.pos)))

    def grad_neg_log_dens(self, state):
        return _grad(self._nld_t)(torch.as_tensor(state.pos)).numpy()

    def log_det_sqrt_gram(self, state):
        return 0.5 * np.linalg.slogdet(self.gram(state))[1]

    def grad_log_det_sqrt_gram(self, state):
        if "grad_logdet" not in state.cache:

            def f(q):
                J = jacfwd(self._constr_t)(q)
                return 0.5 * torch.linalg.slogdet(J @ J.T)[1]

            state.cache["grad_logdet"] = _grad(f)(torch.as_tensor(state.pos)).numpy()
        return state.cache["grad_logdet"]

    def h1(self, state):
        return self.neg_log_dens(state) + self.log_det_sqrt_gram(state)

    def dh1_dpos(self, state):
        return self.grad_neg_log_dens(state) + self.grad_log_det_sqrt_gram(state)

    def h2(self, state):
        return 0.5 * state.mom @ state.mom

    def h(self, state):
        return self.h1(state) + self.h2(state)

    def h1_flow(self, state, dt):
        state.mom = state.mom - dt * self.dh1_dpos(state)

    def h2_flow(self, state, dt):
        state.pos = state.pos + dt * state.mom

    def project_onto_cotangent_space(self, mom, state):
        J = self.jacob_constr(state)
        return mom - J.T @ np.linalg.solve(J @ J.T, J @ mom)

    def sample_momentum(self, state, rng):
        return self.project_onto_cotangent_space(
            rng.standard_normal(state.pos.shape), state
        )


def mici_solve_projection_onto_manifold_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm,
):  # fmt: skip
    """Mici's own Newton solver (used by the toy configs, toy_2d_model.py:170-174).

    Differs from the mlift loops: convergence is tested BEFORE the update is applied,
    so the returned position omits the last computed correction.
    """
    mu = np.zeros_like(state.pos)
    J_prev = system.jacob_constr(state_prev)
    abs_dt = abs(dt)
    for i in range(max_iters):
        J, c = system.jacob_constr(state), system.constr(state)
        error = float(norm(c))
        try:
            delta_mu = J_prev.T @ np.linalg.solve(J @ (abs_dt * J_prev.T), c)
        except np.linalg.LinAlgError as e:
            raise ConvergenceError(f"Newton solver linear algebra failure: {e}")
        delta_pos = abs_dt * delta_mu
        if error > divergence_tol or math.isnan(error):
            raise ConvergenceError(f"Newton solver diverged on iteration {i}.")
        if error < constraint_tol and float(norm(delta_pos)) < position_tol:
            if state.mom is not None:
                state.mom = state.mom - np.sign(dt) * mu
            state.last_num_iters = i
            return state
        mu += delta_mu
        state.pos = state.pos - delta_pos
    raise ConvergenceError(f"Newton solver did not converge in {max_iters} iterations.")


def mici_solve_projection_onto_manifold_quasi_newton(
    state, state_prev, dt, system, constraint_tol=1e-9, position_tol=1e-8,
    divergence_tol=1e10, max_iters=50, norm=maximum_norm,
):  # fmt: skip
    """Mici's symmetric quasi-Newton solver (fixed Gram of the previous point)."""
    mu = np.zeros_like(state.pos)
    J_prev = system.jacob_constr(state_prev)
    abs_dt = abs(dt)
    G_prev = J_prev @ (abs_dt * J_prev.T)
    for i in range(max_iters):
        c = system.constr(state)
        error = float(norm(c))
        delta_mu = J_prev.T @ np.linalg.solve(G_prev, c)
        delta_pos = abs_dt * delta_mu
        if error > divergence_tol or math.isnan(error):
            raise ConvergenceError(f"Quasi-Newton solver diverged on iteration {i}.")
        if error < constraint_tol and float(norm(delta_pos)) < position_tol:
            if state.mom is not None:
                state.mom = state.mom - np.sign(dt) * mu
            state.last_num_iters = i
            return state
        mu += delta_mu
        state.pos = state.pos - delta_pos
    raise ConvergenceError(
        f"Quasi-Newton solver did not converge in {max_iters} iterations."
    )


class ConstrainedLeapfrogIntegrator:
    """A(dt/2) B(dt) A(dt/2) with projection and reversibility check (Appendix A)."""

    def __init__(
        self, system, step_size=None, n_inner_step=1, reverse_check_tol=2e-8,
        reverse_check_norm=maximum_norm,
        projection_solver=mici_solve_projection_onto_manifold_newton,
        projection_solver_kwargs=None,
    ):  # fmt: skip
        self.system, self.step_size, self.n_inner_step = system, step_size, n_inner_step
        self.reverse_check_tol, self.reverse_check_norm = reverse_check_tol, reverse_check_norm
        self.projection_solver = projection_solver
        self.projection_solver_kwargs = (
            {} if projection_solver_kwargs is None else projection_solver_kwargs
        )
        self.last_iters = (0, 0)  # (forward, reverse) of the most recent step

    def _retract(self, state, state_prev, dt):
        self.system.h2_flow(state, dt)
        self.projection_solver(
            state, state_prev, dt, self.system, **self.projection_solver_kwargs
        )

    def _step_a(self, state, dt):
        self.system.h1_flow(state, dt)
        state.mom = self.system.project_onto_cotangent_space(state.mom, state)

    def _step_b(self, state, dt):
        dt_i = dt / self.n_inner_step
        for i in range(self.n_inner_step):
            state_prev = state.copy()
            self._retract(state, state_prev, dt_i)
            it_fwd = state.last_num_iters
            if i == self.n_inner_step - 1:
                self.system.dh1_dpos(state)  # pre-evaluate (caches J, Gram, gradient)
            state.mom = self.system.project_onto_cotangent_space(state.mom, state)
            state_back = state.copy()
            self._retract(state_back, state, -dt_i)
            self.last_iters = (it_fwd, state_back.last_num_iters)
            rev_diff = float(self.reverse_check_norm(state_back.pos - state_prev.pos))
            if rev_diff > self.reverse_check_tol:
                raise NonReversibleStepError(
                    f"Non-reversible step. Distance between initial and "
                    f"forward-backward integrated positions = {rev_diff:.1e}."
                )

    def step(self, state):
        state = state.copy()
        dt = state.dir * self.step_size
        self._step_a(state, 0.5 * dt)
        self._step_b(state, dt)
        self._step_a(state, 0.5 * dt)
        return state


def static_metropolis_hmc_transition(
    system, integrator, state, n_step, mom_noise, unif, max_delta_h=1000.0
):
    """One StaticMetropolisHMC transition driven by supplied random numbers.

    `mom_noise` ~ N(0, I) (projected here as system.sample_momentum does) and `unif`
    ~ U(0,1) are passed in so that a device implementation can be checked decision by
    decision.  Returns (new_state, stats).
    """
    state = state.copy()
    state.mom = system.project_onto_cotangent_space(np.array(mom_noise), state)
    h_init = system.h(state)
    stats = dict(convergence_error=False, non_reversible_step=False, diverging=False,
                 accept_stat=0.0, n_step=0, accepted=False)  # fmt: skip
    prop = state
    try:
        for s in range(n_step):
            prop = integrator.step(prop)
            stats["n_step"] = s + 1
            if abs(system.h(prop) - h_init) > max_delta_h:
                stats["diverging"] = True
                break
    except ConvergenceError:
        stats["convergence_error"] = True
    except NonReversibleStepError:
        stats["non_reversible_step"] = True
    failed = (
        stats["convergence_error"] or stats["non_reversible_step"] or stats["diverging"]
    )
    if not failed:
        h_final = system.h(prop)
        ratio = math.exp(min(h_init - h_final, 0.0)) if not math.isnan(h_final) else 0.0
        stats["accept_stat"] = ratio
        if unif < ratio:
            prop.dir = state.dir
            state, stats["accepted"] = prop, True
    return state, stats


class DualAveragingStepSizeAdapter:
    """Nesterov dual averaging of log step size (Mici defaults; utils.py:281-285)."""

    def __init__(self, adapt_stat_target=0.8, log_step_size_reg_coefficient=0.05,
                 iter_decay_coeff=0.75, iter_offset=10):  # fmt: skip
        self.target, self.reg = adapt_stat_target, log_step_size_reg_coefficient
        self.decay, self.offset = iter_decay_coeff, iter_offset

    def initialize(self, step_size):
        return dict(iter=0, smoothed_log_step_size=0.0, adapt_stat_error=0.0,
                    log_step_size_reg_target=math.log(10 * step_size))  # fmt: skip

    def update(self, a, accept_stat):
        a["iter"] += 1
        w = 1 / (self.offset + a["iter"])
        a["adapt_stat_error"] = (1 - w) * a["adapt_stat_error"] + w * (
            self.target - accept_stat
        )
        sw = (1 / a["iter"]) ** self.decay
        log_eps = a["log_step_size_reg_target"] - (
            a["adapt_stat_error"] * a["iter"] ** 0.5 / self.reg
        )
        a["smoothed_log_step_size"] = (1 - sw) * a["smoothed_log_step_size"] + sw * log_eps
        return math.exp(log_eps)

    @staticmethod
    def finalize(adapt_states):
        return sum(math.exp(a["smoothed_log_step_size"]) for a in adapt_states) / len(
            adapt_states
        )
