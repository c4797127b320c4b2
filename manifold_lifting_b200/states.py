"""Batched chain state (the Mici `ChainState` the reference's solvers mutate,
solvers.py:81-85; created at utils.py:644-647), holding all chains of this rank at once.

`pos` / `mom` are fp64 CUDA tensors of shape [n_chain, dim_q] in the component-major
layout of the C ABI.  Quantities that depend only on `pos` are memoised in `cache`
(cf. mici `cache_in_state`, systems.py:369-405) and dropped when `pos` is reassigned.
"""
import torch

from . import _lib


class ChainState:
    def __init__(self, pos, mom=None, dir=1, _call_counts=None):
        self._pos = _lib.to_soa(pos)
        self.mom = None if mom is None else _lib.to_soa(mom)
        self.dir = dir
        self.cache = {}
        self._call_counts = _call_counts
        self.status = None  # per-chain int32 status of the last solver / integrator call
        self.last_num_iters = None

    @property
    def pos(self):
        return self._pos

    @pos.setter
    def pos(self, value):
        self._pos = _lib.to_soa(value)
        self.cache = {}

    @property
    def n_chain(self):
        return self._pos.shape[0]

    def copy(self):
        new = ChainState.__new__(ChainState)
        new._pos = _lib.soa_clone(self._pos)
        new.mom = None if self.mom is None else _lib.soa_clone(self.mom)
        new.dir = self.dir
        new.cache = dict(self.cache)
        new._call_counts = self._call_counts
        new.status = self.status
        new.last_num_iters = self.last_num_iters
        return new

    def select(self, mask):
        """Sub-state of the chains where `mask` is true (copies)."""
        idx = torch.nonzero(mask).ravel()
        return ChainState(self._pos[idx], None if self.mom is None else self.mom[idx],
                          self.dir, self._call_counts)  # fmt: skip
