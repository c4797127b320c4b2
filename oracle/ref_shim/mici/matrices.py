class IdentityMatrix:
    """Implicit identity (only constructed by mlift/systems.py:428-429, never applied)."""

    def __init__(self, size=None):
        self.size = size

    def __rmul__(self, other):
        return self

    def __matmul__(self, other):
        return other

    def __rmatmul__(self, other):
        return other
