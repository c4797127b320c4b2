"""Error types mirroring `mici.errors` as used by the reference (solvers.py:3,86-95;
adapters.py:2,68)."""


class Error(Exception):
    pass


class IntegratorError(Error, RuntimeError):
    pass


class ConvergenceError(IntegratorError):
    pass


class NonReversibleStepError(IntegratorError):
    pass


class HamiltonianDivergenceError(IntegratorError):
    pass


class AdaptationError(Error, RuntimeError):
    pass
