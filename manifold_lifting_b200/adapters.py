"""Adapters - drop-in for `mlift.adapters` (reference mlift/adapters.py:5-72) plus the
Mici `DualAveragingStepSizeAdapter` the reference pairs it with (utils.py:281-285).

`ToleranceAdapter` keeps the reference's constructor and initialize / update / finalize
protocol and still works by mutating `integrator.projection_solver_kwargs` and
`integrator.reverse_check_tol`, because libmlb takes tolerances per call.
"""
import math

import torch

from . import _lib
from .errors import AdaptationError, IntegratorError


class ToleranceAdapter:
    """Coarser solver tolerances in warm-up, tight ones for the main phase."""

    is_fast = True

    def __init__(self, warm_up_constraint_tol=1e-6, warm_up_position_tol=1e-5,
                 warm_up_reverse_check_tol=None, main_constraint_tol=1e-9,
                 main_position_tol=1e-8, main_reverse_check_tol=None):  # fmt: skip
        self.warm_up_constraint_tol = warm_up_constraint_tol
        self.warm_up_position_tol = warm_up_position_tol
        self.warm_up_reverse_check_tol = (2 * warm_up_position_tol
                                          if warm_up_reverse_check_tol is None
                                          else warm_up_reverse_check_tol)  # fmt: skip
        self.main_constraint_tol = main_constraint_tol
        self.main_position_tol = main_position_tol
        self.main_reverse_check_tol = (2 * main_position_tol
                                       if main_reverse_check_tol is None
                                       else main_reverse_check_tol)  # fmt: skip

    def initialize(self, chain_state, transition):
        kw = transition.integrator.projection_solver_kwargs
        kw["constraint_tol"] = self.warm_up_constraint_tol
        kw["position_tol"] = self.warm_up_position_tol
        transition.integrator.reverse_check_tol = self.warm_up_reverse_check_tol
        return {}

    def update(self, adapt_state, chain_state, trans_stats, transition):
        pass

    def finalize(self, adapt_states, chain_states, transition, rngs=None):
        """Re-project every chain to the main tolerances (adapters.py:52-72).  With a
        batched state the projection of all chains is one launch; any failing chain
        raises `AdaptationError` as in the reference."""
        integrator, system = transition.integrator, transition.system
        kw = integrator.projection_solver_kwargs
        kw["constraint_tol"] = self.main_constraint_tol
        kw["position_tol"] = self.main_position_tol
        states = chain_states if isinstance(chain_states, (list, tuple)) else [chain_states]
        for state in states:
            try:
                integrator.projection_solver(state, state, integrator.step_size_scalar(),
                                             system, raise_on_failure=True, **kw)  # fmt: skip
            except IntegratorError as e:
                raise AdaptationError(
                    "Could not project within specified tolerances for main phase."
                ) from e
        integrator.reverse_check_tol = self.main_reverse_check_tol


class DualAveragingStepSizeAdapter:
    """Per-chain Nesterov dual averaging of log step size, state held on the device
    ([ADAPT_ROWS, n_chain], rows documented in include/mlb.h) and updated INSIDE the
    sampling kernel after every transition; `finalize` sets the integrator step size to
    the mean over chains of exp(smoothed log step size) - over ALL ranks when
    torch.distributed is initialised (one scalar all-reduce; the only collective the
    warm-up needs)."""

    is_fast = True

    def __init__(self, adapt_stat_target=0.8, log_step_size_reg_coefficient=0.05,
                 iter_decay_coeff=0.75, iter_offset=10):  # fmt: skip
        self.adapt_stat_target = adapt_stat_target
        self.log_step_size_reg_coefficient = log_step_size_reg_coefficient
        self.iter_decay_coeff, self.iter_offset = iter_decay_coeff, iter_offset

    def opts(self):
        return _lib.AdaptOpts(self.adapt_stat_target, self.log_step_size_reg_coefficient,
                              self.iter_decay_coeff, self.iter_offset)  # fmt: skip

    max_init_step_size_iters = 100

    def find_init_step_size(self, chain_state, system, integrator, rng=None):
        """Mici 0.2.1 `DualAveragingStepSizeAdapter._find_and_set_init_step_size`, run for
        every chain at once (the reference harness constructs the integrator without a
        step size, example_models/utils.py:317-327, and relies on this search): start from
        step size 1, take ONE step from the initial state, and halve (if
        |h_init - h(step)| > log 2, is NaN or the step raised an integrator error) or
        double the step size until the comparison with log 2 flips.  Returns the per-chain
        step sizes [n_chain]; chains that never flip within `max_init_step_size_iters`
        raise `AdaptationError` as Mici does."""
        from .errors import AdaptationError

        state = chain_state.copy()
        if getattr(state, "mom", None) is None:
            gen = rng if rng is not None else torch.Generator(device="cuda").manual_seed(0)
            state.mom = system.sample_momentum(state, gen)
        n = state.n_chain
        eps = torch.ones(n, dtype=torch.float64, device="cuda")
        too_big = torch.zeros(n, dtype=torch.bool, device="cuda")
        done = torch.zeros(n, dtype=torch.bool, device="cuda")
        log2 = math.log(2.0)
        for s in range(self.max_init_step_size_iters):
            out = integrator.steps(state, 1, step_size=eps, raise_on_failure=False)
            if s == 0 and bool(torch.isnan(out.h_init).any()):
                raise AdaptationError("Hamiltonian evaluating to NaN at initial state.")
            failed = out.status != 0
            delta_h = (out.h_init - out.h_final).abs()
            nan = torch.isnan(delta_h) & ~failed
            over = delta_h > log2
            if s == 0:
                too_big = nan | over
            too_big = torch.where(nan | failed, torch.ones_like(too_big), too_big)
            flipped = ~failed & ~nan & ((too_big & ~over) | (~too_big & over))
            done |= flipped
            if bool(done.all()):
                return eps
            scale = torch.where(too_big, 0.5, 2.0).to(torch.float64)
            eps = torch.where(done, eps, eps * scale)
        raise AdaptationError(
            f"Could not find reasonable initial step size in {self.max_init_step_size_iters} "
            f"iterations for {int((~done).sum())} of {n} chains. A very large final step "
            "size may indicate that the target distribution is improper.")

    def initialize(self, chain_state, transition):
        n = chain_state.n_chain
        integrator = transition.integrator
        a = torch.zeros(_lib.ADAPT_ROWS, n, dtype=torch.float64, device="cuda")
        if integrator.step_size is None:
            eps = self.find_init_step_size(chain_state, transition.system, integrator)
            integrator.step_size = float(eps.mean())
            a[3] = torch.log(10 * eps)
            a[4] = eps
        else:
            eps = float(integrator.step_size_scalar())
            a[3] = math.log(10 * eps)
            a[4] = eps
        return a

    def finalize(self, adapt_state, chain_states, transition, rngs=None):
        tot = torch.stack([torch.exp(adapt_state[1]).sum(),
                           torch.tensor(float(adapt_state.shape[1]), dtype=torch.float64,
                                        device=adapt_state.device)])  # fmt: skip
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(tot)
        transition.integrator.step_size = float(tot[0] / tot[1])


__all__ = ["ToleranceAdapter", "DualAveragingStepSizeAdapter"]
