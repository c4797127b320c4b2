"""`jax.experimental.ode` stand-in: only importable; the adaptive integrator is outside
the constrained-HMC hot path (mlift/ode.py:56-) and is not restated."""


def odeint(*args, **kwargs):
    raise NotImplementedError("adaptive odeint is outside the pinned path")
