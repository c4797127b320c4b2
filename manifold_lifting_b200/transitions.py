"""`mlift.transitions` (reference mlift/transitions.py:9,66,129) holds position-dependent
random-walk Metropolis / MALA baselines used only by the acceptance-grid experiment
(example_models/acceptance_statistics_grid.py:97-126).  They are NOT part of constrained
HMC and are out of scope for the B200 hot path; the module is
kept importable so that `import mlift.transitions` does not break callers."""


def _out_of_scope(name):
    class _Stub:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"{name} is a CPU baseline transition of the reference's acceptance-grid "
                "experiment and is outside the scope of the B200 constrained-HMC hot path"
            )

    _Stub.__name__ = name
    return _Stub


PositionDependentRWMTransition = _out_of_scope("PositionDependentRWMTransition")
SimplePositionDependentMALATransition = _out_of_scope("SimplePositionDependentMALATransition")
XifaraPositionDependentMALATransition = _out_of_scope("XifaraPositionDependentMALATransition")
