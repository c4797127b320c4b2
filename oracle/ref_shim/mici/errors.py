class Error(Exception):
    pass


class IntegratorError(Error):
    pass


class NonReversibleStepError(IntegratorError):
    pass


class ConvergenceError(IntegratorError):
    pass


class AdaptationError(Error):
    pass


class LinAlgError(Error):
    pass
