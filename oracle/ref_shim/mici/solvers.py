"""Placeholder module (`mici.solvers` is referenced inside example_models/utils.py:297-303
only); Mici's own solvers are restated in oracle/mlift_oracle/mici.py."""


def solve_projection_onto_manifold_newton(*a, **k):
    raise NotImplementedError


def solve_projection_onto_manifold_quasi_newton(*a, **k):
    raise NotImplementedError


def maximum_norm(vct):
    return abs(vct).max()
