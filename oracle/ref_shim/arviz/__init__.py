"""Placeholder `arviz` (imported at module level by example_models/utils.py:12; its
functions are only called by the experiment harness, which is out of scope)."""


def summary(*a, **k):
    raise NotImplementedError("arviz is not installable here")


rhat = ess = summary
