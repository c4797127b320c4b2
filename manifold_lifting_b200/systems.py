"""Constrained systems for differentiable generative models - batched CUDA drop-in for
`mlift.systems` (reference mlift/systems.py:148-807).

Same class names, method names and private "jitted" attributes as the reference
(`_constr ... _newton_projection_with_line_search`, systems.py:328-346, which
`mlift.solvers` and `example_models/utils.py:432-433` touch), but every method acts on
ALL chains of a `ChainState` at once: arrays gain a leading chain axis
(`pos` is [n_chain, dim_q], `dy_du` is [n_chain, dim_y, dim_u], ...).  All arithmetic runs
in hand-written sm_100a kernels behind the C ABI of include/mlb.h; there is no CPU
fallback.

`generate_y` must be a `manifold_lifting_b200.models.ModelSpec` (see models.py for why a
Python closure cannot be accepted).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, p_soa, p_vec, soa_empty, stream_ptr, to_soa
from .models import ModelSpec
from .solvers import euclidean_norm, maximum_norm


def _norm_code(norm):
    if norm is maximum_norm or norm == "maximum" or norm == _lib.NORM_MAX:
        return _lib.NORM_MAX
    if norm is euclidean_norm or norm == "euclidean" or norm == _lib.NORM_EUCLIDEAN:
        return _lib.NORM_EUCLIDEAN
    raise ValueError("norm must be mlift.solvers.maximum_norm or euclidean_norm "
                     "(a Python callable cannot run inside the projection kernel)")  # fmt: skip


class _Packed(tuple):
    """Tuple of block views that remembers the packed [n_chain, rows] device array."""

    def __new__(cls, views, packed):
        self = super().__new__(cls, views)
        self.packed = packed
        return self


class JacobBlocks(_Packed):
    """(dy_du, dy_dn) or (dy_du, dy_dv, dy_dn) - systems.py:580, 710."""


class GramComponents(_Packed):
    """(chol_gram,) [dense] or (chol_cap_mtx, d) [Woodbury] - systems.py:594, 617, 757."""


def standard_normal_neg_log_dens(q):
    """systems.py:26-28 (kept for API compatibility; NumPy / torch input)."""
    return 0.5 * (q**2).sum(-1)


def standard_normal_grad_neg_log_dens(q):
    """systems.py:31-33."""
    return q, 0.5 * (q**2).sum(-1)


def convert_to_numpy_pytree(tree):
    """systems.py:36-49 analogue: device tensors -> NumPy arrays, recursively."""
    if isinstance(tree, torch.Tensor):
        return tree.detach().cpu().numpy()
    if isinstance(tree, (tuple, list)):
        return type(tree)(convert_to_numpy_pytree(t) for t in tree) if not isinstance(
            tree, _Packed) else tuple(convert_to_numpy_pytree(t) for t in tree)  # fmt: skip
    if isinstance(tree, dict):
        return {k: convert_to_numpy_pytree(v) for k, v in tree.items()}
    return tree


def construct_mici_system_neg_log_dens_functions(neg_log_dens):
    """systems.py:52-87.  The extended-prior density of a compiled-in model lives in the
    CUDA library; this returns markers telling the system class to use it."""
    return "model", "model"


class _AbstractDifferentiableGenerativeModelSystem:
    """Matrix-free constrained system (systems.py:148-471), batched over chains."""

    hierarchical = False

    def __init__(self, model, neg_log_dens=None, grad_neg_log_dens=None):
        if not isinstance(model, ModelSpec):
            raise TypeError(
                "generate_y must be a manifold_lifting_b200.models.ModelSpec: an arbitrary "
                "Python closure cannot cross the C ABI into the CUDA kernels"
            )
        for f in (neg_log_dens, grad_neg_log_dens):
            if not (f is None or f == "model" or f is standard_normal_neg_log_dens
                    or f is standard_normal_grad_neg_log_dens):  # fmt: skip
                raise TypeError("custom neg_log_dens callables are not supported: the "
                                "model's extended-prior density is compiled into libmlb")  # fmt: skip
        _lib.require_cuda()
        self.model = model
        self.dim_u, self.dim_y, self.dim_q = model.dim_u, model.dim_y, model.dim_q
        self._h = model.handle
        self._L = lib()

    # ------------------------------------------------------------------ packing
    def _jac_views(self, packed):
        U, Y = self.dim_u, self.dim_y
        n = packed.shape[0]
        dy_du = packed[:, : Y * U].unflatten(1, (Y, U))
        if self.hierarchical:
            views = (dy_du, packed[:, Y * U : Y * U + Y], packed[:, Y * U + Y :])
        else:
            views = (dy_du, packed[:, Y * U :])
        assert views[-1].shape == (n, Y)
        return JacobBlocks(views, packed)

    def _gram_views(self, packed):
        K = self.model.gram_dim
        chol = packed[:, : K * K].unflatten(1, (K, K))
        d = packed[:, K * K :]
        views = (chol,) if K == self.dim_y else (chol, d)
        return GramComponents(views, packed)

    def _pack_jac(self, blocks):
        if isinstance(blocks, JacobBlocks):
            return blocks.packed
        blocks = [torch.as_tensor(b, dtype=torch.float64, device="cuda") for b in blocks]
        n = blocks[0].shape[0] if blocks[0].ndim == 3 else 1
        flat = [b.reshape(n, -1) for b in blocks]
        return to_soa(torch.cat(flat, 1))

    def _pack_gram(self, comps):
        if isinstance(comps, GramComponents):
            return comps.packed
        comps = [torch.as_tensor(c, dtype=torch.float64, device="cuda") for c in comps]
        n = comps[0].shape[0] if comps[0].ndim == 3 else 1
        flat = [c.reshape(n, -1) for c in comps]
        if len(flat) == 1:  # dense branch: d is not part of the reference's tuple
            flat.append(torch.ones(n, self.dim_y, dtype=torch.float64, device="cuda"))
        return to_soa(torch.cat(flat, 1))

    def _split_blocks(self, args, n_sets=1):
        """Split `(*jac_blocks[, *jac_blocks_2], vct)` positional args as the reference
        passes them (systems.py:357-361)."""
        per = 3 if self.hierarchical else 2
        sets = [tuple(args[i * per : (i + 1) * per]) for i in range(n_sets)]
        return sets, args[n_sets * per :]

    # ------------------------------------- array-level callables (systems.py:328-346)
    def _constr(self, q):
        q = to_soa(q)
        n = q.shape[0]
        c = soa_empty(n, self.dim_y)
        (pq, ld), (pc, _) = p_soa(q, self.dim_q), p_soa(c)
        check(self._L.mlb_constr(self._h, C.c_int64(n), C.c_int64(ld), pq, pc, stream_ptr()))
        return c

    def _jacob_constr_blocks(self, q):
        q = to_soa(q)
        n = q.shape[0]
        jac, c = soa_empty(n, self.model.jac_rows), soa_empty(n, self.dim_y)
        (pq, ld) = p_soa(q, self.dim_q)
        check(self._L.mlb_jacob_constr_blocks(self._h, C.c_int64(n), C.c_int64(ld), pq,
                                              p_soa(jac)[0], p_soa(c)[0], stream_ptr()))  # fmt: skip
        return self._jac_views(jac), c

    def _decompose_gram(self, *jac_blocks):
        jac = self._pack_jac(jac_blocks[0] if len(jac_blocks) == 1 else jac_blocks)
        n = jac.shape[0]
        gram = soa_empty(n, self.model.gram_rows)
        pj, ld = p_soa(jac, self.model.jac_rows)
        check(self._L.mlb_decompose_gram(self._h, C.c_int64(n), C.c_int64(ld), pj,
                                         p_soa(gram)[0], stream_ptr()))  # fmt: skip
        return self._gram_views(gram)

    def _logdet_common(self, q, with_grad):
        q = to_soa(q)
        n = q.shape[0]
        val = torch.empty(n, dtype=torch.float64, device="cuda")
        c = soa_empty(n, self.dim_y)
        jac, gram = soa_empty(n, self.model.jac_rows), soa_empty(n, self.model.gram_rows)
        pq, ld = p_soa(q, self.dim_q)
        if with_grad:
            grad = soa_empty(n, self.dim_q)
            check(self._L.mlb_val_and_grad_log_det_sqrt_gram(
                self._h, C.c_int64(n), C.c_int64(ld), pq, p_vec(val), p_soa(grad)[0],
                p_soa(c)[0], p_soa(jac)[0], p_soa(gram)[0], stream_ptr()))  # fmt: skip
        else:
            grad = None
            check(self._L.mlb_log_det_sqrt_gram(
                self._h, C.c_int64(n), C.c_int64(ld), pq, p_vec(val), p_soa(c)[0],
                p_soa(jac)[0], p_soa(gram)[0], stream_ptr()))  # fmt: skip
        return val, (c, self._jac_views(jac), self._gram_views(gram)), grad

    def _log_det_sqrt_gram(self, q):
        val, aux, _ = self._logdet_common(q, False)
        return val, aux

    def _val_and_grad_log_det_sqrt_gram(self, q):
        val, aux, grad = self._logdet_common(q, True)
        return (val, aux), grad

    def _block_op(self, fn, jac, vct, rows_in, rows_out, gram=None, jac_r=None):
        jac = self._pack_jac(jac)
        vct = to_soa(vct)
        n = vct.shape[0]
        out = soa_empty(n, rows_out)
        pj, ld = p_soa(jac, self.model.jac_rows)
        args = [self._h, C.c_int64(n), C.c_int64(ld), pj]
        if jac_r is not None:
            args.append(p_soa(self._pack_jac(jac_r), self.model.jac_rows)[0])
        if gram is not None:
            args.append(p_soa(self._pack_gram(gram), self.model.gram_rows)[0])
        args += [p_soa(vct, rows_in)[0], p_soa(out)[0], stream_ptr()]
        check(fn(*args))
        return out

    def _lmult_by_jacob_constr(self, *args):
        (jac,), (vct,) = self._split_blocks(args) if len(args) > 2 else ((args[0],), (args[1],))
        return self._block_op(self._L.mlb_lmult_by_jacob_constr, jac, vct, self.dim_q,
                              self.dim_y)  # fmt: skip

    def _rmult_by_jacob_constr(self, *args):
        (jac,), (vct,) = self._split_blocks(args) if len(args) > 2 else ((args[0],), (args[1],))
        return self._block_op(self._L.mlb_rmult_by_jacob_constr, jac, vct, self.dim_y,
                              self.dim_q)  # fmt: skip

    def _lmult_by_pinv_jacob_constr(self, jacob_constr_blocks, gram_components, vct):
        return self._block_op(self._L.mlb_lmult_by_pinv_jacob_constr, jacob_constr_blocks,
                              vct, self.dim_y, self.dim_q, gram=gram_components)  # fmt: skip

    def _lmult_by_inv_jacob_product(self, *args):
        if len(args) == 3:
            jac_l, jac_r, vct = args
        else:
            (jac_l, jac_r), (vct,) = self._split_blocks(args, 2)
        return self._block_op(self._L.mlb_lmult_by_inv_jacob_product, jac_l, vct, self.dim_y,
                              self.dim_y, jac_r=jac_r)  # fmt: skip

    def _normal_space_component(self, jacob_constr_blocks, gram_components, vct):
        return self._block_op(self._L.mlb_normal_space_component, jacob_constr_blocks, vct,
                              self.dim_q, self.dim_q, gram=gram_components)  # fmt: skip

    def _solve(self, solver, q, jac_prev, gram_prev, dt, constraint_tol, position_tol,
               divergence_tol, max_iters, norm, max_line_search_iters=10):  # fmt: skip
        q = to_soa(q)
        n = q.shape[0]
        opts = _lib.SolverOpts(solver, int(max_iters), int(max_line_search_iters),
                               _norm_code(norm), constraint_tol, position_tol,
                               divergence_tol, 2 * position_tol, _lib.NORM_MAX, 0)  # fmt: skip
        q_out, mu = soa_empty(n, self.dim_q), soa_empty(n, self.dim_q)
        iters, n_constr, status = (torch.empty(n, dtype=torch.int32, device="cuda")
                                   for _ in range(3))  # fmt: skip
        ndq, err = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2))
        pq, ld = p_soa(q, self.dim_q)
        pj = p_soa(self._pack_jac(jac_prev), self.model.jac_rows)[0]
        pg = None if gram_prev is None else p_soa(self._pack_gram(gram_prev),
                                                  self.model.gram_rows)[0]  # fmt: skip
        check(self._L.mlb_solve_projection(
            self._h, C.c_int64(n), C.c_int64(ld), pq, pj, pg, C.c_double(dt), C.byref(opts),
            p_soa(q_out)[0], p_soa(mu)[0], p_vec(iters, torch.int32), p_vec(ndq), p_vec(err),
            p_vec(n_constr, torch.int32), p_vec(status, torch.int32), stream_ptr()))  # fmt: skip
        return q_out, mu, iters, ndq, err, n_constr, status

    def _quasi_newton_projection(self, q, jacob_constr_blocks_prev, gram_components_prev, dt,
                                 constraint_tol, position_tol, divergence_tol, max_iters,
                                 norm=maximum_norm):  # fmt: skip
        """systems.py:190-228.  Returns (q, mu/dt, i, norm_delta_q, error) per chain."""
        r = self._solve(_lib.SOLVER_QUASI_NEWTON, q, jacob_constr_blocks_prev,
                        gram_components_prev, dt, constraint_tol, position_tol,
                        divergence_tol, max_iters, norm)  # fmt: skip
        return r[0], r[1], r[2], r[3], r[4]

    def _newton_projection(self, q, jacob_constr_blocks_prev, dt, constraint_tol,
                           position_tol, divergence_tol, max_iters, norm=maximum_norm):  # fmt: skip
        """systems.py:230-270."""
        r = self._solve(_lib.SOLVER_NEWTON, q, jacob_constr_blocks_prev, None, dt,
                        constraint_tol, position_tol, divergence_tol, max_iters, norm)  # fmt: skip
        return r[0], r[1], r[2], r[3], r[4]

    def _newton_projection_with_line_search(self, q, jacob_constr_blocks_prev, dt,
                                            constraint_tol, position_tol, divergence_tol,
                                            max_iters, max_line_search_iters,
                                            norm=maximum_norm):  # fmt: skip
        """systems.py:272-326.  Returns (..., num_constr_calls) as sixth element."""
        r = self._solve(_lib.SOLVER_NEWTON_LS, q, jacob_constr_blocks_prev, None, dt,
                        constraint_tol, position_tol, divergence_tol, max_iters, norm,
                        max_line_search_iters)  # fmt: skip
        return r[0], r[1], r[2], r[3], r[4], r[5]

    def precompile_jax_functions(self, q, solver_norm=maximum_norm):
        """systems.py:349-367.  Nothing to compile (kernels are built ahead of time);
        runs every callable once, which also warms the CUDA context."""
        q = to_soa(q)
        jac, c = self._jacob_constr_blocks(q)
        gram = self._decompose_gram(jac)
        self._constr(q), self._log_det_sqrt_gram(q), self._val_and_grad_log_det_sqrt_gram(q)
        self._lmult_by_jacob_constr(jac, q), self._rmult_by_jacob_constr(jac, c)
        self._lmult_by_pinv_jacob_constr(jac, gram, c)
        self._lmult_by_inv_jacob_product(jac, jac, c)
        self._normal_space_component(jac, gram, q)
        self._quasi_newton_projection(q, jac, gram, 1.0, 0.1, 0.1, 1.0, 10, solver_norm)
        self._newton_projection(q, jac, 1.0, 0.1, 0.1, 1.0, 10, solver_norm)
        self._newton_projection_with_line_search(q, jac, 1.0, 0.1, 0.1, 1.0, 10, 10,
                                                 solver_norm)  # fmt: skip

    # --------------------------------- Mici-facing cached methods (systems.py:369-471)
    def _count(self, state, name, k=1):
        if state._call_counts is not None:
            state._call_counts[name] = state._call_counts.get(name, 0) + k

    def constr(self, state):
        if "constr" not in state.cache:
            state.cache["constr"] = self._constr(state.pos)
            self._count(state, "constr")
        return state.cache["constr"]

    def jacob_constr_blocks(self, state):
        if "jac" not in state.cache:
            state.cache["jac"], state.cache["constr"] = self._jacob_constr_blocks(state.pos)
            self._count(state, "jacob_constr_blocks")
        return state.cache["jac"]

    def gram_components(self, state):
        if "gram" not in state.cache:
            state.cache["gram"] = self._decompose_gram(self.jacob_constr_blocks(state))
            self._count(state, "gram_components")
        return state.cache["gram"]

    def log_det_sqrt_gram(self, state):
        if "logdet" not in state.cache:
            val, (c, jac, gram) = self._log_det_sqrt_gram(state.pos)
            state.cache.update(logdet=val, constr=c, jac=jac, gram=gram)
            self._count(state, "log_det_sqrt_gram")
        return state.cache["logdet"]

    def grad_log_det_sqrt_gram(self, state):
        if "grad_logdet" not in state.cache:
            (val, (c, jac, gram)), grad = self._val_and_grad_log_det_sqrt_gram(state.pos)
            state.cache.update(grad_logdet=grad, logdet=val, constr=c, jac=jac, gram=gram)
            self._count(state, "grad_log_det_sqrt_gram")
        return state.cache["grad_logdet"]

    def _nld(self, state):
        if "nld" not in state.cache:
            q = state.pos
            n = q.shape[0]
            val = torch.empty(n, dtype=torch.float64, device="cuda")
            grad = soa_empty(n, self.dim_q)
            pq, ld = p_soa(q, self.dim_q)
            check(self._L.mlb_neg_log_dens(self._h, C.c_int64(n), C.c_int64(ld), pq,
                                           p_vec(val), p_soa(grad)[0], stream_ptr()))  # fmt: skip
            state.cache["nld"], state.cache["grad_nld"] = val, grad
            self._count(state, "neg_log_dens")
        return state.cache["nld"], state.cache["grad_nld"]

    def neg_log_dens(self, state):
        return self._nld(state)[0]

    def grad_neg_log_dens(self, state):
        return self._nld(state)[1]

    def h1(self, state):
        return self.neg_log_dens(state) + self.log_det_sqrt_gram(state)

    def dh1_dpos(self, state):
        return self.grad_neg_log_dens(state) + self.grad_log_det_sqrt_gram(state)

    def h2(self, state):
        return 0.5 * (state.mom * state.mom).sum(-1)

    def h(self, state):
        return self.h1(state) + self.h2(state)

    def dh2_dmom(self, state):
        return state.mom

    def dh2_dpos(self, state):
        return 0 * state.pos

    def dh_dpos(self, state):
        return self.dh1_dpos(state)

    def h1_flow(self, state, dt):
        """Mici System.h1_flow: mom -= dt * dh1_dpos."""
        state.mom = to_soa(state.mom - dt * self.dh1_dpos(state))

    def h2_flow(self, state, dt):
        state.pos = to_soa(state.pos + dt * self.dh2_dmom(state))

    def normal_space_component(self, state, vct):
        return self._normal_space_component(self.jacob_constr_blocks(state),
                                            self.gram_components(state), vct)  # fmt: skip

    def lmult_by_jacob_constr(self, state, vct):
        return self._lmult_by_jacob_constr(self.jacob_constr_blocks(state), vct)

    def rmult_by_jacob_constr(self, state, vct):
        return self._rmult_by_jacob_constr(self.jacob_constr_blocks(state), vct)

    def lmult_by_pinv_jacob_constr(self, state, vct):
        return self._lmult_by_pinv_jacob_constr(self.jacob_constr_blocks(state),
                                                self.gram_components(state), vct)  # fmt: skip

    def lmult_by_inv_jacob_product(self, state_1, state_2, vct):
        return self._lmult_by_inv_jacob_product(self.jacob_constr_blocks(state_1),
                                                self.jacob_constr_blocks(state_2), vct)  # fmt: skip

    def project_onto_cotangent_space(self, mom, state):
        mom = to_soa(mom)
        mom -= self.normal_space_component(state, mom)
        return mom

    def sample_momentum(self, state, rng):
        """systems.py:468-471.  `rng`: numpy Generator (host draw, as the reference) or
        torch.Generator (device draw)."""
        shape = tuple(state.pos.shape)
        if isinstance(rng, torch.Generator):
            mom = to_soa(torch.randn(shape, dtype=torch.float64, device="cuda", generator=rng))
        else:
            mom = to_soa(np.asarray(rng.standard_normal(shape)))
        return self.project_onto_cotangent_space(mom, state)


class IndependentAdditiveNoiseModelSystem(_AbstractDifferentiableGenerativeModelSystem):
    """y = F(u) + s(u) T(n) (systems.py:474-654).  Constructor signature as the reference;
    `generate_y` is a ModelSpec of the additive-noise family."""

    hierarchical = False

    def __init__(self, generate_y, data=None, dim_u=None, neg_log_dens=None,
                 grad_neg_log_dens=None, force_full_gram_cholesky=False,
                 jacob_generate_y=None):  # fmt: skip
        super().__init__(generate_y, neg_log_dens, grad_neg_log_dens)
        if generate_y.family != "additive":
            raise TypeError("model is not of the independent-additive-noise family")
        if dim_u is not None and dim_u != self.dim_u:
            raise ValueError(f"dim_u={dim_u} does not match the model ({self.dim_u})")
        dense = self.dim_u >= self.dim_y
        if force_full_gram_cholesky and not dense:
            raise NotImplementedError("force_full_gram_cholesky: only the branch the "
                                      "reference would pick by default is compiled")  # fmt: skip
        if jacob_generate_y is not None:
            raise TypeError("jacob_generate_y closures cannot cross the C ABI")
        self.data = data if data is not None else generate_y.data


class HierarchicalLatentVariableModelSystem(_AbstractDifferentiableGenerativeModelSystem):
    """y = g(u, v, n) (systems.py:657-807); `generate_y` is a hierarchical ModelSpec."""

    hierarchical = True

    def __init__(self, generate_y, data=None, dim_u=None, neg_log_dens=None,
                 grad_neg_log_dens=None):  # fmt: skip
        super().__init__(generate_y, neg_log_dens, grad_neg_log_dens)
        if generate_y.family != "hierarchical":
            raise TypeError("model is not of the hierarchical latent variable family")
        if dim_u is not None and dim_u != self.dim_u:
            raise ValueError(f"dim_u={dim_u} does not match the model ({self.dim_u})")
        self.data = data if data is not None else generate_y.data


__all__ = [
    "standard_normal_neg_log_dens", "standard_normal_grad_neg_log_dens",
    "convert_to_numpy_pytree", "construct_mici_system_neg_log_dens_functions",
    "IndependentAdditiveNoiseModelSystem", "HierarchicalLatentVariableModelSystem",
    "JacobBlocks", "GramComponents",
]  # fmt: skip
