"""`jax.config` stand-in: fp64 on CPU is the only mode, `update` records and ignores."""


class _Config:
    def __init__(self):
        self.values = {}

    def update(self, key, value):
        self.values[key] = value


config = _Config()
update = config.update
