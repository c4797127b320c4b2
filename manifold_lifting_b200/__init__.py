"""manifold_lifting_b200 - B200-native constrained-HMC hot path of `mlift`.

Exports mirror `mlift/__init__.py:1-3` of the reference (systems, adapters, solvers);
`install_as_mlift()` additionally aliases this package as `mlift` so reference-style
imports (`import mlift`, `from mlift.systems import ...`) resolve to the CUDA path.
"""
import sys

from . import (_lib, adapters, errors, gaussian_process, integrators, linalg, models, outputs,
               parallel, samplers, solvers, state_space, states, systems, transitions)  # fmt: skip
from .adapters import *  # noqa: F401,F403
from .solvers import *  # noqa: F401,F403
from .systems import *  # noqa: F401,F403


def install_as_mlift():
    """Register this package under the name `mlift` (drop-in for the reference)."""
    me = sys.modules[__name__]
    sys.modules.setdefault("mlift", me)
    for sub in ("systems", "solvers", "adapters", "transitions", "linalg"):
        sys.modules.setdefault(f"mlift.{sub}", getattr(me, sub))
    return me
