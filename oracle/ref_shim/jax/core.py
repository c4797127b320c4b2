class ConcretizationTypeError(TypeError):
    """Never raised here: nothing is traced, every value is concrete."""
