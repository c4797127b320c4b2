"""Minimal `mici` stand-in (Mici 0.2.1 is pinned by the reference's environment.yml:17 and
is NOT in /root/reference nor installable here).

TEST INFRASTRUCTURE ONLY (see oracle/ref_shim/README.md).  Two kinds of content:

  * what the reference's modules IMPORT and subclass — `systems.System`, `states.ChainState`
    + `cache_in_state*`, `errors`, `adapters.Adapter`, `matrices.IdentityMatrix` — restated
    from Mici's published behaviour as constrained by the reference's call sites
    (mlift/systems.py:369-471, solvers.py:56-95, adapters.py:1-72, utils.py:645);
  * `integrators.ConstrainedLeapfrogIntegrator`, RESTATED (not executed) from SURVEY.md
    Appendix A, so that a whole constrained leapfrog step can be driven through the
    reference's own system classes and solver wrappers.
"""
from . import adapters, errors, integrators, matrices, samplers, solvers, states, systems  # noqa: F401
