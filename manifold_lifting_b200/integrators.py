"""Constrained leapfrog integrator - the Mici `ConstrainedLeapfrogIntegrator` the
reference constructs at mlift/example_models/utils.py:317-327 and
toy_2d_model.py:183-185, batched over chains.

Two execution paths with identical semantics (tests check one against the other):
  * `step` / `steps` - ONE fused kernel per call (`mlb_leapfrog_steps`): A(dt/2), B(dt)
    with projection + reversibility check, A(dt/2); chain state stays in registers for
    all `n_step` steps.  This is the hot path.
  * `step_unfused` - the same step composed on the host from the individual system /
    solver calls, in the order Mici issues them.  This is what a caller
    that drives the system method by method gets; it exists to prove the boundary is a
    true drop-in and for debugging.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, p_soa, p_vec, stream_ptr
from .errors import ConvergenceError, NonReversibleStepError
from .solvers import (SOLVER_IDS, jitted_solve_projection_onto_manifold_newton,
                      maximum_norm)  # fmt: skip
from .systems import _norm_code


class ConstrainedLeapfrogIntegrator:
    def __init__(self, system, step_size=None, n_inner_step=1, reverse_check_tol=2e-8,
                 reverse_check_norm=maximum_norm,
                 projection_solver=jitted_solve_projection_onto_manifold_newton,
                 projection_solver_kwargs=None):  # fmt: skip
        if n_inner_step != 1:
            raise NotImplementedError("n_inner_step > 1 is not used by the reference")
        self.system, self.step_size, self.n_inner_step = system, step_size, n_inner_step
        self.reverse_check_tol, self.reverse_check_norm = reverse_check_tol, reverse_check_norm
        # ODE-constrained models only (include/mlb.h MLB_FLAG_SINGLE_KERNEL): True selects
        # the asynchronous one-thread-per-chain kernel instead of the host-driven
        # sub-phase kernels; both compute the same step
        self.single_kernel = False
        self.projection_solver = projection_solver
        # mutated in place by ToleranceAdapter (adapters.py:45-47, 56-58)
        self.projection_solver_kwargs = ({} if projection_solver_kwargs is None
                                         else projection_solver_kwargs)  # fmt: skip

    def solver_opts(self):
        kw = self.projection_solver_kwargs
        try:
            solver = SOLVER_IDS[self.projection_solver]
        except KeyError:
            raise TypeError("projection_solver must be one of the solvers in "
                            "manifold_lifting_b200.solvers") from None  # fmt: skip
        return _lib.SolverOpts(
            solver, int(kw.get("max_iters", 50)), int(kw.get("max_line_search_iters", 10)),
            _norm_code(kw.get("norm", maximum_norm)), kw.get("constraint_tol", 1e-9),
            kw.get("position_tol", 1e-8), kw.get("divergence_tol", 1e10),
            self.reverse_check_tol, _norm_code(self.reverse_check_norm),
            _lib.FLAG_SINGLE_KERNEL if self.single_kernel else 0)  # fmt: skip

    def step_size_scalar(self):
        ss = self.step_size
        return float(ss.mean()) if isinstance(ss, torch.Tensor) else float(ss)

    # ------------------------------------------------------------------ fused path
    def steps(self, state, n_step=1, step_size=None, raise_on_failure=None, energies=True):
        """`n_step` constrained leapfrog steps for every chain in one launch.  Returns a
        new state; `state.status[i] != 0` marks chains whose step failed (they stay at
        the last completed step, `state.n_done[i]` steps done).  `energies=False` skips
        the Hamiltonian at the two ends of the trajectory (`h_init` / `h_final` stay
        None): Mici's integrator does not form it either, only its samplers do, and its
        logarithms are 16 % of a 32-step eight-schools launch."""
        sysm = self.system
        new = state.copy()
        n = new.n_chain
        dt = self.step_size if step_size is None else step_size
        per_chain = isinstance(dt, torch.Tensor)
        dt_vec = (new.dir * dt).contiguous() if per_chain else None
        dt_scalar = 0.0 if per_chain else float(new.dir * dt)
        status, n_done, itf, itr = (torch.empty(n, dtype=torch.int32, device="cuda")
                                    for _ in range(4))  # fmt: skip
        h0, h1 = ((torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2))
                  if energies else (None, None))  # fmt: skip
        opts = self.solver_opts()
        pq, ld = p_soa(new.pos, sysm.dim_q)
        pp, ld2 = p_soa(new.mom, sysm.dim_q)
        assert ld == ld2
        check(lib().mlb_leapfrog_steps(
            sysm._h, C.c_int64(n), C.c_int64(ld), pq, pp, C.c_double(dt_scalar),
            p_vec(dt_vec), C.c_int32(n_step), C.byref(opts), p_vec(status, torch.int32),
            p_vec(n_done, torch.int32), p_vec(itf, torch.int32), p_vec(itr, torch.int32),
            p_vec(h0), p_vec(h1), stream_ptr()))  # fmt: skip
        new.cache = {}
        new.status, new.n_done = status, n_done
        new.iters_fwd, new.iters_rev, new.h_init, new.h_final = itf, itr, h0, h1
        if raise_on_failure or (raise_on_failure is None and n == 1):
            st = int(status[0]) if n == 1 else int(status.max())
            if st == _lib.ST_NON_REVERSIBLE:
                raise NonReversibleStepError("Non-reversible step.")
            if st != _lib.ST_OK:
                raise ConvergenceError(f"Projection solver failed (status {st}).")
        return new

    def step(self, state):
        """Mici `Integrator.step`: one step of size `state.dir * step_size`."""
        return self.steps(state, 1)

    # ----------------------------------------------------------------- unfused path
    def _retract(self, state, state_prev, dt):
        self.system.h2_flow(state, dt)
        self.projection_solver(state, state_prev, dt, self.system, raise_on_failure=True,
                               **self.projection_solver_kwargs)  # fmt: skip

    def _step_a(self, state, dt):
        self.system.h1_flow(state, dt)
        state.mom = self.system.project_onto_cotangent_space(state.mom, state)

    def step_unfused(self, state):
        """The same step, operation by operation (raises on any chain's failure)."""
        state = state.copy()
        dt = state.dir * self.step_size
        self._step_a(state, 0.5 * dt)
        state_prev = state.copy()
        self._retract(state, state_prev, dt)
        iters_fwd = state.last_num_iters
        self.system.dh1_dpos(state)
        state.mom = self.system.project_onto_cotangent_space(state.mom, state)
        state_back = state.copy()
        self._retract(state_back, state, -dt)
        rev = self.reverse_check_norm(state_back.pos - state_prev.pos)
        if bool((rev > self.reverse_check_tol).any()):
            raise NonReversibleStepError(
                f"Non-reversible step. Distance between initial and forward-backward "
                f"integrated positions = {float(rev.max()):.1e}.")  # fmt: skip
        self._step_a(state, 0.5 * dt)
        state.iters_fwd, state.iters_rev = iters_fwd, state_back.last_num_iters
        return state
