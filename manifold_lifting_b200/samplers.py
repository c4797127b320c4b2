"""Static-trajectory Metropolis HMC over all chains, resident on the GPU.

Replaces, for the constrained path, the Mici sampler loop the reference drives from
`example_models/utils.py:438-453` (`sampler.sample_chains`) with ONE kernel launch per
phase: every chain runs its whole sequence of transitions (momentum refresh ->
`n_step` constrained leapfrog steps -> Metropolis accept, semantics of the notebook's
`simulate_proposal`, cell 44, i.e. Mici `StaticMetropolisHMC`) with its state in
registers, writing only traces.  Dual-averaging step-size adaptation runs inside the
kernel per chain; the cross-chain mean at the end of warm-up is the only reduction.

Random numbers are counter-based (Philox4x32-10 keyed by seed, GLOBAL chain index,
iteration), so a run is bit-identical however the chains are sharded over GPUs.
`DynamicMultinomialHMC` is the reference's default sampler (utils.py:286-288): same
driver, with the tree built per chain inside the kernel (csrc/mlb_nuts.cuh).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, p_soa, p_vec, stream_ptr
from .adapters import DualAveragingStepSizeAdapter
from .states import ChainState


class StaticMetropolisHMC:
    def __init__(self, system, integrator, rng=None, n_step=1, seed=None,
                 max_delta_h=1000.0):  # fmt: skip
        self.system, self.integrator, self.n_step = system, integrator, n_step
        self.max_delta_h = max_delta_h
        if seed is None:
            seed = 0 if rng is None else int(rng.integers(0, 2**63 - 1))
        self.seed = int(seed)
        self.iteration = 0  # global transition counter (Philox key component)

    # `transition`-like attribute access used by the adapters
    @property
    def transition(self):
        return self

    def _launch(self, q, n_iter, adapt=None, adapt_opts=None, trace_q=None, trace_every=1,
                trace_accept=None, stats=None, last_status=None, chain_offset=0,
                mom_noise=None, unif=None, trace_dim=0):  # fmt: skip
        sysm, integ = self.system, self.integrator
        n = q.shape[0]
        opts = integ.solver_opts()
        pq, ld = p_soa(q, sysm.dim_q)
        if integ.step_size is None:
            raise ValueError(
                "integrator.step_size is None: pass a step size or run a warm-up phase with "
                "a DualAveragingStepSizeAdapter (which searches for an initial one)")
        # a per-chain tensor of step sizes enters through its mean, as in the dynamic
        # sampler (per-chain values only live in the adapter's state during warm-up)
        check(lib().mlb_hmc_sample(
            sysm._h, C.c_int64(n), C.c_int64(ld), pq,
            C.c_double(integ.step_size_scalar()),
            C.c_int32(self.n_step), C.c_int32(n_iter), C.byref(opts),
            C.c_double(self.max_delta_h), C.c_uint64(self.seed), C.c_uint64(chain_offset),
            C.c_uint32(self.iteration), p_vec(mom_noise), p_vec(unif), p_vec(adapt),
            None if adapt_opts is None else C.byref(adapt_opts), p_vec(trace_q),
            C.c_int32(trace_every), C.c_int32(trace_dim), p_vec(trace_accept), p_vec(stats),
            p_vec(last_status, torch.int32), stream_ptr()))  # fmt: skip
        self.iteration += n_iter

    def sample_chains(self, n_warm_up_iter, n_main_iter, init_states, adapters=(),
                      trace_funcs=None, trace_every=1, chain_offset=0,
                      max_iters_per_launch=None, trace_dim=None, **unused):  # fmt: skip
        """Returns (final_state, traces, stats).

        traces[name] has shape [n_chain, n_main_iter // trace_every, ...] (`trace_dim`
        limits the "pos" trace to the leading components, e.g. dim_u); stats holds
        per-chain means of accept_stat and failure indicators over the main phase
        (names as monitored by the reference, utils.py:335-340)."""
        state = init_states if isinstance(init_states, ChainState) else ChainState(init_states)
        q = _lib.soa_clone(state.pos)
        n, Q = q.shape
        step_adapter = next((a for a in adapters
                             if isinstance(a, DualAveragingStepSizeAdapter)), None)  # fmt: skip
        other = [a for a in adapters if a is not step_adapter]
        if n_warm_up_iter > 0:
            cs = ChainState(q)
            for a in other:
                a.initialize(cs, self)
            adapt = step_adapter.initialize(cs, self) if step_adapter else None
            wstats = torch.zeros(_lib.STAT_ROWS, n, dtype=torch.float64, device="cuda")
            self._launch(q, n_warm_up_iter, adapt=adapt,
                         adapt_opts=step_adapter.opts() if step_adapter else None,
                         stats=wstats, chain_offset=chain_offset)  # fmt: skip
            if step_adapter:
                step_adapter.finalize(adapt, None, self)
            cs = ChainState(q)
            for a in other:
                a.finalize(None, cs, self)
            q = _lib.soa_clone(cs.pos)
            self.warm_up_stats = wstats
        n_trace = n_main_iter // trace_every
        stats = torch.zeros(_lib.STAT_ROWS, n, dtype=torch.float64, device="cuda")
        last = torch.zeros(n, dtype=torch.int32, device="cuda")
        tdim = Q if not trace_dim else int(trace_dim)
        trace_q = torch.empty(max(n_trace, 1), tdim, n, dtype=torch.float64, device="cuda")
        chunk = max_iters_per_launch or n_main_iter
        chunk = max(trace_every, (chunk // trace_every) * trace_every)
        done = 0
        while done < n_main_iter:
            k = min(chunk, n_main_iter - done)
            tq = trace_q[done // trace_every :] if n_trace > 0 else None
            self._launch(q, k, trace_q=tq if (tq is not None and tq.shape[0] > 0) else None,
                         trace_every=trace_every, stats=stats, last_status=last,
                         chain_offset=chain_offset, trace_dim=tdim)  # fmt: skip
            done += k
        pos_trace = trace_q[:n_trace].permute(2, 0, 1)  # [n_chain, n_trace, Q] (view)
        traces = {"pos": pos_trace}
        for f in trace_funcs or ():
            traces.update(f(pos_trace))
        n_it = max(n_main_iter, 1)
        out_stats = {
            "accept_stat": stats[0] / n_it,
            "accepted": stats[1] / n_it,
            "convergence_error": stats[2] / n_it,
            "non_reversible_step": stats[3] / n_it,
            "diverging": stats[4] / n_it,
            "n_step": stats[5],
            "projection_iters": stats[6],
            "tree_depth": stats[7] / n_it,
            "last_status": last,
        }
        return ChainState(q), traces, out_stats


class DynamicMultinomialHMC(StaticMetropolisHMC):
    """Mici `DynamicMultinomialHMC(system, integrator, rng, max_tree_depth=10,
    max_delta_h=1000)` as the reference constructs it (example_models/utils.py:286-288),
    batched: every chain builds its own tree inside one kernel (toy / hierarchical models)
    or all trees grow in host-driven lock-step rounds (ODE models) (`mlb_nuts_sample`).
    `sample_chains` and the adapters work as for the static sampler; `stats` additionally
    reports the mean tree depth, and `n_step` the integrator steps taken."""

    def __init__(self, system, integrator, rng=None, max_tree_depth=10, max_delta_h=1000.0,
                 do_extra_subtree_checks=True, seed=None, two_sided_max_delta_h=False):  # fmt: skip
        super().__init__(system, integrator, rng=rng, n_step=0, seed=seed,
                         max_delta_h=max_delta_h)  # fmt: skip
        self.max_tree_depth = int(max_tree_depth)
        # Mici 0.2.1 defaults: no-U-turn criterion also on the overlapping sub-trajectories
        # of every merge, divergence when h - h_init > max_delta_h (one-sided)
        self.do_extra_subtree_checks = bool(do_extra_subtree_checks)
        self.two_sided_max_delta_h = bool(two_sided_max_delta_h)
        self.trace_n_step = self.trace_depth = None  # set to int32 [n_iter, n] tensors to record

    def _launch(self, q, n_iter, adapt=None, adapt_opts=None, trace_q=None, trace_every=1,
                trace_accept=None, stats=None, last_status=None, chain_offset=0,
                mom_noise=None, unif=None, trace_dim=0):  # fmt: skip
        if unif is not None:
            raise ValueError("the dynamic sampler draws its uniforms from Philox")
        sysm, integ = self.system, self.integrator
        n = q.shape[0]
        opts = integ.solver_opts()
        if not self.do_extra_subtree_checks:
            opts.flags |= _lib.FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS
        if self.two_sided_max_delta_h:
            opts.flags |= _lib.FLAG_NUTS_TWO_SIDED_DELTA_H
        pq, ld = p_soa(q, sysm.dim_q)
        check(lib().mlb_nuts_sample(
            sysm._h, C.c_int64(n), C.c_int64(ld), pq, C.c_double(integ.step_size_scalar()),
            C.c_int32(self.max_tree_depth), C.c_int32(n_iter), C.byref(opts),
            C.c_double(self.max_delta_h), C.c_uint64(self.seed), C.c_uint64(chain_offset),
            C.c_uint32(self.iteration), p_vec(mom_noise), p_vec(adapt),
            None if adapt_opts is None else C.byref(adapt_opts), p_vec(trace_q),
            C.c_int32(trace_every), C.c_int32(trace_dim), p_vec(trace_accept),
            p_vec(self.trace_n_step, torch.int32), p_vec(self.trace_depth, torch.int32),
            p_vec(stats), p_vec(last_status, torch.int32), stream_ptr()))  # fmt: skip
        self.iteration += n_iter
