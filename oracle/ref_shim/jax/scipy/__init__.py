from . import linalg, special  # noqa: F401
