"""`mici.systems.System` base: what mlift's system classes inherit (mlift/systems.py:148)."""
from .states import cache_in_state, cache_in_state_with_aux


class System:
    def __init__(self, neg_log_dens, grad_neg_log_dens=None):
        self._neg_log_dens = neg_log_dens
        self._grad_neg_log_dens = grad_neg_log_dens

    @cache_in_state("pos")
    def neg_log_dens(self, state):
        return self._neg_log_dens(state.pos)

    @cache_in_state_with_aux("pos", "neg_log_dens")
    def grad_neg_log_dens(self, state):
        return self._grad_neg_log_dens(state.pos)

    def h1(self, state):
        return self.neg_log_dens(state)

    def dh1_dpos(self, state):
        return self.grad_neg_log_dens(state)

    def h1_flow(self, state, dt):
        state.mom -= dt * self.dh1_dpos(state)

    def h(self, state):
        return self.h1(state) + self.h2(state)


class EuclideanMetricSystem(System):
    """Placeholder: only named as a default argument (example_models/utils.py:531)."""


class ConstrainedEuclideanMetricSystem(EuclideanMetricSystem):
    """Placeholder: only named in an issubclass test (example_models/utils.py:297)."""


class DenseConstrainedEuclideanMetricSystem(ConstrainedEuclideanMetricSystem):
    """Placeholder (the toy script's system; restated in oracle/mlift_oracle/mici.py)."""
