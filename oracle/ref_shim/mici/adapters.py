class Adapter:
    """Interface mlift.adapters.ToleranceAdapter implements (mlift/adapters.py:5-72)."""

    is_fast = True

    def initialize(self, chain_state, transition):
        raise NotImplementedError

    def update(self, adapt_state, chain_state, trans_stats, transition):
        raise NotImplementedError

    def finalize(self, adapt_states, chain_states, transition, rngs):
        raise NotImplementedError
