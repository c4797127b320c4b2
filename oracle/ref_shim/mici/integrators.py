"""Constrained leapfrog integrator, RESTATED from Mici 0.2.1's published behaviour
 - Mici's source is not available here, so this file is the one part
of a reference-driven step that is not the reference's / its dependency's own code.

`_step = A(dt/2) B(dt) A(dt/2)`; A: `h1_flow` then cotangent projection; B per inner step:
copy -> `h2_flow` -> projection solver -> (last inner step) pre-evaluate `dh1_dpos` ->
cotangent projection -> reverse check through a second projection solve.
"""
from .errors import NonReversibleStepError
from .solvers import maximum_norm


class ConstrainedLeapfrogIntegrator:
    def __init__(self, system, step_size=None, n_inner_step=1, reverse_check_tol=2e-8,
                 reverse_check_norm=maximum_norm, projection_solver=None,
                 projection_solver_kwargs=None):  # fmt: skip
        self.system, self.step_size, self.n_inner_step = system, step_size, n_inner_step
        self.reverse_check_tol = reverse_check_tol
        self.reverse_check_norm = reverse_check_norm
        self.projection_solver = projection_solver
        self.projection_solver_kwargs = (
            {} if projection_solver_kwargs is None else projection_solver_kwargs
        )

    def _h2_flow_retraction(self, state, state_prev, dt):
        self.system.h2_flow(state, dt)
        self.projection_solver(state, state_prev, dt, self.system,
                               **self.projection_solver_kwargs)  # fmt: skip

    def _project_onto_cotangent_space(self, state):
        state.mom = self.system.project_onto_cotangent_space(state.mom, state)

    def _step_a(self, state, dt):
        self.system.h1_flow(state, dt)
        self._project_onto_cotangent_space(state)

    def _step_b(self, state, dt):
        dt_i = dt / self.n_inner_step
        for i in range(self.n_inner_step):
            state_prev = state.copy()
            self._h2_flow_retraction(state, state_prev, dt_i)
            if i == self.n_inner_step - 1:
                self.system.dh1_dpos(state)
            self._project_onto_cotangent_space(state)
            state_back = state.copy()
            self._h2_flow_retraction(state_back, state, -dt_i)
            rev_diff = self.reverse_check_norm(state_back.pos - state_prev.pos)
            if rev_diff > self.reverse_check_tol:
                raise NonReversibleStepError(
                    f"Non-reversible step. Distance between initial and "
                    f"forward-backward integrated positions = {rev_diff:.1e}.")

    def step(self, state):
        state = state.copy()
        dt = state.dir * self.step_size
        self._step_a(state, 0.5 * dt)
        self._step_b(state, dt)
        self._step_a(state, 0.5 * dt)
        return state
