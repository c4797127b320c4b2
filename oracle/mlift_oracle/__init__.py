"""CPU oracle for the constrained-HMC hot path of mlift (TEST INFRASTRUCTURE ONLY).

This package is a torch-fp64 / NumPy restatement of the reference algorithm
(`/root/reference/mlift/systems.py`, `solvers.py`, `ode.py`, `prior.py`, the four
BASELINE model files) and of the Mici 0.2.1 integrator / solver / sampler semantics the
reference relies on (Mici's source is NOT in the reference tree: restated from its
public behaviour).

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for the
constrained path and neither JAX nor Mici can be installed in this container or on the
GPU box, so the oracle cannot be checked against the reference's own outputs. It is
pinned only by (i) the Euclidean-leapfrog known-answer vector printed in the reference
notebook, (ii) two independent restatements checking one another (this autodiff
restatement vs the closed-form C restatement in `oracle/c`), and (iii) the invariants
listed in SURVEY.md section 8c.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import
this package.  The product (`manifold_lifting_b200`) never does.
"""
import torch

torch.set_default_dtype(torch.float64)

from . import prior, ode, models, systems, solvers, mici  # noqa: E402,F401
