#!/usr/bin/env python
"""Writes tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

`/root/reference/mlift/{systems,solvers,linalg,ode,prior,distributions,transforms,math}.py`
and `example_models/{eight_schools,fitzhugh_nagumo,hh_current_stimulus,
stochastic_volatility}.py` are imported unmodified (oracle/ref_loader.py) on top of the
`jax` / `mici` stand-ins of oracle/ref_shim (torch-fp64 eager arithmetic; JAX and Mici are
not installable in this image).  Every array below is therefore an output of the
reference's code, not of a restatement:

  * the 13 jitted callables of `mlift/systems.py:328-346` at a handful of states,
  * the three projection loops (`systems.py:190-326`) called directly and through the
    solver wrappers (`solvers.py:16-256`), including their failure branches,
  * whole constrained leapfrog steps: the reference's system + solver wrappers driven by
    the Mici integrator as RESTATED in oracle/ref_shim/mici/integrators.py (Mici's source is
    not available; that file is the only non-reference code on the path),
  * `mlift.linalg.tridiagonal_solve / tridiagonal_pos_def_log_det` and the closures of
    `PartiallyInvertibleStateSpaceModelSystem` on the stochastic-volatility model.

The system objects are built exactly as `example_models/utils.py:292-327, 572-580` builds
them.  This script needs /root/reference, so it runs in the build container only; the
fixtures it writes are committed and travel to the GPU box.

    python tests/golden/make_ref_golden.py [hier] [fn] [hh] [sv] [linalg] [dense]
"""
import os
import sys
import time
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle")]

import ref_loader  # noqa: E402

mlift = ref_loader.load()
import mici  # noqa: E402  (the stand-in)

DATA = os.path.join(ref_loader.REFERENCE_ROOT, "data")
SOURCE = ("reference source executed through oracle/ref_shim (torch-fp64 jax stand-in); "
          "integrator step ordering restated (Mici 0.2.1 not available)")  # fmt: skip
SOLVERS = {
    0: ("quasi_newton", mlift.jitted_solve_projection_onto_manifold_quasi_newton),
    1: ("newton", mlift.jitted_solve_projection_onto_manifold_newton),
    2: ("newton_line_search",
        mlift.jitted_solve_projection_onto_manifold_newton_with_line_search),
}  # fmt: skip
TOLS = dict(constraint_tol=1e-9, position_tol=1e-8, divergence_tol=1e10, max_iters=50)
ST_OK, ST_DIVERGED, ST_NOT_CONVERGED, ST_NON_REVERSIBLE = 0, 1, 2, 3


def build_system(cls, module, data, **kw):
    """example_models/utils.py:572-580 + :296."""
    nld, gnld = mlift.construct_mici_system_neg_log_dens_functions(
        partial(module.extended_prior_neg_log_dens, data=data))  # fmt: skip
    return cls(data=data, dim_u=module.compute_dim_u(data), neg_log_dens=nld,
               grad_neg_log_dens=gnld, **kw)  # fmt: skip


def state_of(q, p=None):
    return mici.states.ChainState(
        pos=np.array(q, dtype=np.float64), mom=None if p is None else np.array(p),
        dir=1, _call_counts={})  # fmt: skip


def flat_blocks(jb):
    """(dy_du, [dy_dv,] dy_dn) -> dict of arrays."""
    if len(jb) == 2:
        return dict(dy_du=jb[0], dy_dn=jb[1])
    return dict(dy_du=jb[0], dy_dv=jb[1], dy_dn=jb[2])


def record_ops(system, qs, q_prevs, rng, hier):
    """The 13 callables at each state (blocks of `q_prevs` as the 'previous' Jacobian)."""
    out = {}

    def put(k, v):
        out.setdefault(k, []).append(np.asarray(v, dtype=np.float64))

    for q, qp in zip(qs, q_prevs):
        st, stp = state_of(q), state_of(qp)
        Q = q.size
        jb = system.jacob_constr_blocks(st)
        Y = jb[-1].size
        vq, wy = rng.standard_normal(Q), rng.standard_normal(Y)
        put("q", q), put("q_prev", qp), put("vec_q", vq), put("vec_y", wy)
        put("c", system.constr(st))
        for k, v in flat_blocks(jb).items():
            put(k, v)
        gc = system.gram_components(st)
        put("chol", gc[0]), put("d", gc[1])
        put("logdet", system.log_det_sqrt_gram(state_of(q)))
        put("grad_logdet", system.grad_log_det_sqrt_gram(st))
        put("nld", system.neg_log_dens(st)), put("grad_nld", system.grad_neg_log_dens(st))
        put("h1", system.h1(st)), put("dh1_dpos", system.dh1_dpos(st))
        put("jv", system.lmult_by_jacob_constr(st, vq))
        put("jtw", system.rmult_by_jacob_constr(st, wy))
        put("pinv_w", system.lmult_by_pinv_jacob_constr(st, wy))
        put("nsc_v", system.normal_space_component(st, vq))
        put("ijp_w", system.lmult_by_inv_jacob_product(st, stp, wy))
        put("proj_v", system.project_onto_cotangent_space(vq.copy(), st))
    return {f"ops_{k}": np.array(v) for k, v in out.items()}


def record_loops(system, qs, ps, dt, norm_names=("max",), max_ls=10):
    """The three loops called directly (`system._quasi_newton_projection` ...), from
    q + dt p with the blocks / Gram factor of q as the previous point."""
    out = {"loop_dt": dt}
    norms = {"max": mlift.maximum_norm, "l2": mlift.euclidean_norm}
    for nn in norm_names:
        rec = {}
        for q, p in zip(qs, ps):
            st = state_of(q)
            jb, gc = system.jacob_constr_blocks(st), system.gram_components(st)
            q_start = q + dt * p
            args = (dt, TOLS["constraint_tol"], TOLS["position_tol"],
                    TOLS["divergence_tol"], TOLS["max_iters"])  # fmt: skip
            res = {
                0: system._quasi_newton_projection(q_start, jb, gc, *args, norms[nn]),
                1: system._newton_projection(q_start, jb, *args, norms[nn]),
                2: system._newton_projection_with_line_search(
                    q_start, jb, *args, max_ls, norms[nn]),
            }  # fmt: skip
            rec.setdefault("q_start", []).append(q_start)
            for sid, r in res.items():
                r = mlift.systems.convert_to_numpy_pytree(tuple(r))
                for name, v in zip(("q", "mu", "iters", "norm_dq", "err", "n_constr"), r):
                    rec.setdefault(f"{name}_{sid}", []).append(np.asarray(v, np.float64))
        for k, v in rec.items():
            out[f"loop_{nn}_{k}"] = np.array(v)
    out["loop_q_prev"], out["loop_p"] = np.array(qs), np.array(ps)
    return out


class IterationLog:
    """Records the iteration count every loop call returns (third output)."""

    def __init__(self, system):
        self.log = []
        for name in ("_quasi_newton_projection", "_newton_projection",
                     "_newton_projection_with_line_search"):  # fmt: skip
            setattr(system, name, self._wrap(getattr(system, name)))

    def _wrap(self, fn):
        def f(*a):
            r = fn(*a)
            self.log.append(int(r[2]))
            return r

        return f


def record_steps(system, itlog, qs, ps, dt, n_step, solver_ids=(0, 1, 2), tag="step",
                 tols=None):  # fmt: skip
    """`n_step` constrained leapfrog steps per chain and solver; per step: position,
    momentum, Hamiltonian, (forward, reverse) iteration counts; a failing step records
    the status the exception maps to and stops that chain."""
    tols = dict(TOLS if tols is None else tols)
    tols.pop("divergence_tol", None)
    out = {f"{tag}_q0": np.array(qs), f"{tag}_p0": np.array(ps), f"{tag}_dt": dt,
           f"{tag}_n_step": n_step, f"{tag}_solvers": np.array(solver_ids)}  # fmt: skip
    for sid in solver_ids:
        Q = np.array(qs).shape[1]
        n = len(qs)
        q_t = np.full((n, n_step, Q), np.nan)
        p_t = np.full((n, n_step, Q), np.nan)
        h_t = np.full((n, n_step), np.nan)
        it_t = np.full((n, n_step, 2), -1)
        status = np.zeros(n, np.int64)
        n_done = np.zeros(n, np.int64)
        h0 = np.zeros(n)
        for i, (q, p) in enumerate(zip(qs, ps)):
            integ = mici.integrators.ConstrainedLeapfrogIntegrator(
                system, dt, projection_solver=SOLVERS[sid][1],
                reverse_check_tol=2 * tols["position_tol"],
                projection_solver_kwargs=dict(tols))  # fmt: skip
            st = state_of(q, p)
            h0[i] = system.h(st)
            for s in range(n_step):
                del itlog.log[:]
                try:
                    st = integ.step(st)
                except mici.errors.ConvergenceError as e:
                    status[i] = ST_DIVERGED if "diverged" in str(e) else ST_NOT_CONVERGED
                    break
                except mici.errors.NonReversibleStepError:
                    status[i] = ST_NON_REVERSIBLE
                    break
                q_t[i, s], p_t[i, s], h_t[i, s] = st.pos, st.mom, system.h(st)
                it_t[i, s] = itlog.log[:2]
                n_done[i] = s + 1
        out.update({f"{tag}_q_{sid}": q_t, f"{tag}_p_{sid}": p_t, f"{tag}_h_{sid}": h_t,
                    f"{tag}_iters_{sid}": it_t, f"{tag}_status_{sid}": status,
                    f"{tag}_n_done_{sid}": n_done, f"{tag}_h0": h0})  # fmt: skip
    return out


def tangent_momenta(system, qs, rng):
    return [system.sample_momentum(state_of(q), rng) for q in qs]


# ----------------------------------------------------------------------------- models
def hier():
    es = ref_loader.example_model("eight_schools")
    file_data = dict(np.load(os.path.join(DATA, "eight-schools-data.npz")))
    rng16 = np.random.default_rng(202101)
    sig16 = rng16.uniform(9.0, 18.0, 16)
    y16 = 4.4 + 3.6 * rng16.standard_normal(16) + sig16 * rng16.standard_normal(16)
    for name, y_obs, sigma, par, n_prior, n_mod in (
        ("hier_unbounded", file_data["y_obs"], file_data["σ"], "unbounded", 6, 10),
        ("hier_normal", file_data["y_obs"], file_data["σ"], "normal", 6, 10),
        ("hier16_unbounded", y16, sig16, "unbounded", 2, 4),
    ):
        t0 = time.time()
        data = {"y_obs": np.asarray(y_obs, float), "σ": np.asarray(sigma, float),
                "parametrization": par}  # fmt: skip
        system = build_system(mlift.HierarchicalLatentVariableModelSystem, es, data,
                              generate_y=es.generate_y)  # fmt: skip
        itlog = IterationLog(system)
        rng = np.random.default_rng(11)
        Y = data["y_obs"].size
        # (i) the reference's own initial states (prior draws: heavy-tailed tau)
        qs = [np.asarray(q) for q in es.sample_initial_states(rng, data, n_prior)]
        # (ii) moderate latents, lifted as eight_schools.py:75-77 does
        for _ in range(n_mod):
            u = np.array([rng.normal(0, 1), rng.normal(0.5, 0.7)])
            v = rng.standard_normal(Y)
            _, x = es.generate_from_model(u, v, data)
            qs.append(np.concatenate([u, v, (data["y_obs"] - np.asarray(x)) / data["σ"]]))
        ps = tangent_momenta(system, qs, rng)
        off = [q + 0.05 * rng.standard_normal(q.size) for q in qs]  # c != 0
        out = dict(source=SOURCE, y_obs=data["y_obs"], sigma=data["σ"], parametrization=par)
        out.update(record_ops(system, off, qs, rng, hier=True))
        out.update(record_loops(system, qs, ps, 0.15, norm_names=("max", "l2")))
        out.update(record_steps(system, itlog, qs, ps, 0.15, 3))
        # failure branches: a step far too long
        out.update(record_steps(system, itlog, qs[-6:], [4.0 * p for p in ps[-6:]], 2.5, 1,
                                tag="fail"))  # fmt: skip
        np.savez(os.path.join(HERE, f"ref_{name}.npz"), **out)
        print(name, "done in", round(time.time() - t0, 1), "s;",
              {k: out[k].tolist() for k in out if k.startswith(("step_status", "fail_status"))})  # fmt: skip


def _additive(name, module, data, u_centres, dim_u, spread, dt_loop, dt_step, n_ops,
              n_loop, n_step_chains, off_scale, solver_ids=(0, 1, 2)):  # fmt: skip
    t0 = time.time()
    system = build_system(mlift.IndependentAdditiveNoiseModelSystem, module, data,
                          generate_y=module.generate_y)  # fmt: skip
    itlog = IterationLog(system)
    rng = np.random.default_rng(5)

    def lift(u):
        y_mean = np.asarray(module.generate_y(u, np.zeros(data["y_obs"].size), data))
        sigma = float(module.generate_params(u, data)["σ"])
        return np.concatenate([u, (data["y_obs"] - y_mean) / sigma])

    qs = [lift(u_centres + spread * rng.standard_normal(dim_u)) for _ in range(n_ops)]
    out = dict(source=SOURCE, parametrization=data["parametrization"])
    off = [q.copy() for q in qs]
    for q in off:
        q[dim_u:] += off_scale * rng.standard_normal(q.size - dim_u)
    out.update(record_ops(system, off, qs, rng, hier=False))
    print(name, "ops", round(time.time() - t0, 1), "s", flush=True)
    ps = tangent_momenta(system, qs, rng)
    out.update(record_loops(system, qs[:n_loop], ps[:n_loop], dt_loop))
    print(name, "loops", round(time.time() - t0, 1), "s", flush=True)
    out.update(record_steps(system, itlog, qs[:n_step_chains], ps[:n_step_chains], dt_step,
                            1, solver_ids))  # fmt: skip
    np.savez(os.path.join(HERE, f"ref_{name}.npz"), **out)
    print(name, "done in", round(time.time() - t0, 1), "s;",
          {k: out[k].tolist() for k in out if "iters" in k or "status" in k})  # fmt: skip


def fn():
    """FitzHugh-Nagumo exactly as fitzhugh_nagumo.py:146-149 loads it (obs noise 0.1)."""
    m = ref_loader.example_model("fitzhugh_nagumo")
    d = dict(np.load(os.path.join(DATA, "fitzhugh-nagumo-simulated-data.npz")))
    d["y_obs"] = d["y_mean_obs"] + 0.1 * d["n_obs"]
    loc = np.array([0, 0, 0, -1, -3, -2, -1.0, 0, 0])
    u_true = np.concatenate([np.log([3, 1, 3, 1 / 3, 1 / 15, 1 / 15, 0.1]), [-1.0, 1.0]])
    for par, n_ops, n_loop, n_chain in (("unbounded", 3, 2, 2), ("normal", 2, 1, 1)):
        data = dict(d, parametrization=par)
        _additive(f"fn_{par}", m, data, u_true - (loc if par == "normal" else 0), 9, 0.05,
                  0.02, 0.02, n_ops, n_loop, n_chain, 0.01)  # fmt: skip


def hh():
    """Hodgkin-Huxley exactly as hh_current_stimulus.py:287-292 loads it (obs noise 1)."""
    m = ref_loader.example_model("hh_current_stimulus")
    d = dict(np.load(os.path.join(DATA, "hodgkin-huxley-current-stimulus-data.npz")))
    d["y_obs"] = d["x_obs"] + 1.0 * d["n_obs"]
    which = sys.argv[2:] or ["normal", "unbounded"]
    for par in which:
        data = dict(d, parametrization=par)
        u0 = np.asarray(d["u_obs"], float)
        if par == "unbounded":  # same parameters in the other chart
            loc = np.array([8, 4, 2, -3, -3, -60, 0.0])
            scale = np.array([1, 1, 1, 1, 1, 10, 1.0])
            u0 = loc + scale * u0
        _additive(f"hh_{par}", m, data, u0, 7, 0.02 if par == "normal" else 0.0, 0.002,
                  0.002, 1, 1, 1, 0.01, solver_ids=(0, 1) if par == "normal" else (1,))  # fmt: skip


def linalg():
    """mlift/linalg.py:46-114 on random diagonally dominant systems, several sizes."""
    from mlift.linalg import tridiagonal_pos_def_log_det, tridiagonal_solve

    rng = np.random.default_rng(3)
    out = dict(source=SOURCE)
    for k, n in enumerate((2, 3, 17, 200)):
        a, c = rng.standard_normal(n - 1), rng.standard_normal(n - 1)
        b = 2.5 + np.abs(rng.standard_normal(n)) + np.pad(np.abs(a), (1, 0)) + np.pad(np.abs(c), (0, 1))
        d = rng.standard_normal(n)
        out[f"a_{k}"], out[f"b_{k}"], out[f"c_{k}"], out[f"d_{k}"] = a, b, c, d
        out[f"x_{k}"] = np.asarray(tridiagonal_solve(a, b, c, d))
        out[f"logdet_{k}"] = np.asarray(tridiagonal_pos_def_log_det(a, b))
    out["n_case"] = 4
    np.savez(os.path.join(HERE, "ref_linalg.npz"), **out)
    print("linalg done")


def sv():
    """Stochastic-volatility state-space system (stochastic_volatility.py:71-141 through
    PartiallyInvertibleStateSpaceModelSystem, systems.py:1084-1254), dim_y 40, both ways
    of getting the blocks (the model's closed form, and the class's autodiff default)."""
    m = ref_loader.example_model("stochastic_volatility")
    rng = np.random.default_rng(9)
    dim_y = 40
    # synthetic observations from the model itself at plausible parameters
    u_true = np.array([-0.5, np.log(0.3), 2.0])
    data = {"y_obs": np.zeros(dim_y), "parametrization": "unbounded"}
    _, x = m.generate_from_model(u_true, rng.standard_normal(dim_y), data)
    data["y_obs"] = np.exp(np.asarray(x) / 2) * rng.standard_normal(dim_y)
    out = dict(source=SOURCE, y_obs=data["y_obs"])
    for tag, jac in (("closed", m.jacob_constr_split_blocks), ("auto", None)):
        system = build_system(mlift.PartiallyInvertibleStateSpaceModelSystem, m, data,
                              constr_split=m.constr_split, jacob_constr_split_blocks=jac)  # fmt: skip
        r = np.random.default_rng(10)
        rec = {}
        for _ in range(3):
            u = u_true + 0.2 * r.standard_normal(3)
            n = r.standard_normal(dim_y)
            n = np.maximum(np.abs(n), 0.05) * np.sign(data["y_obs"])  # y / n > 0
            v = r.standard_normal(dim_y)
            q, qp = np.concatenate([u, v, n]), np.concatenate([u + 0.01, v, n * 1.01])
            st, stp = state_of(q), state_of(qp)
            (dc_du, dc_dv, dc_dn, dx_dy) = system.jacob_constr_blocks(st)
            a, b, chol = system.gram_components(st)
            vq, wy = r.standard_normal(q.size), r.standard_normal(dim_y)
            vals = dict(q=q, q_prev=qp, vec_q=vq, vec_y=wy, c=system.constr(st), dc_du=dc_du,
                        dc_dv=dc_dv, dc_dn0=dc_dn[0], dc_dn1=dc_dn[1], dx_dy=dx_dy, a=a, b=b,
                        chol=chol, logdet=system.log_det_sqrt_gram(state_of(q)),
                        grad_logdet=system.grad_log_det_sqrt_gram(st),
                        jv=system.lmult_by_jacob_constr(st, vq),
                        jtw=system.rmult_by_jacob_constr(st, wy),
                        pinv_w=system.lmult_by_pinv_jacob_constr(st, wy),
                        nsc_v=system.normal_space_component(st, vq),
                        ijp_w=system.lmult_by_inv_jacob_product(st, stp, wy),
                        nld=system.neg_log_dens(st), grad_nld=system.grad_neg_log_dens(st))  # fmt: skip
            for k, v in vals.items():
                rec.setdefault(k, []).append(np.asarray(v, np.float64))
        out.update({f"{tag}_{k}": np.array(v) for k, v in rec.items()})
        print("sv", tag, "done", flush=True)
    np.savez(os.path.join(HERE, "ref_sv.npz"), **out)


def dense():
    """mlift/linalg.py:6-43 (`lu_triangular_solve`, `jvp_cholesky_mtx_mult_by_vct`) on random
    factors, and the Gaussian-process system built on them (systems.py:810-931) with the
    squared-exponential covariance of gp_regression.py:33-44 on random inputs: blocks,
    Gram factor, both solves, log-determinant and its gradient at a few states."""
    import jax
    from mlift.linalg import jvp_cholesky_mtx_mult_by_vct, lu_triangular_solve

    rng = np.random.default_rng(21)
    out = dict(source=SOURCE)
    dims = (1, 2, 5, 33, 70, 150)
    for k, n in enumerate(dims):
        a, b2 = rng.standard_normal((n, n)), rng.standard_normal((n, n))
        mtx = a @ a.T / n + np.eye(n)
        chol = np.linalg.cholesky(mtx)
        upper = np.linalg.cholesky(b2 @ b2.T / n + np.eye(n)).T
        rhs = rng.standard_normal((n, 4))
        out[f"l_{k}"], out[f"u_{k}"], out[f"b_{k}"] = chol, upper, rhs
        out[f"x_{k}"] = np.asarray(lu_triangular_solve(chol, upper, rhs))
        out[f"xv_{k}"] = np.asarray(lu_triangular_solve(chol, upper, rhs[:, 0]))
        t = rng.standard_normal((n, 3, n))
        t = t + t.transpose(2, 1, 0)  # symmetric in (0, 2) for every direction
        v = rng.standard_normal(n)
        out[f"t_{k}"], out[f"v_{k}"] = t, v
        out[f"jvp_{k}"] = np.asarray(
            jax.vmap(jvp_cholesky_mtx_mult_by_vct, (1, None, None), 1)(t, chol, v))
    out["n_case"] = len(dims)
    # Gaussian-process system on the reference's regression model, synthetic inputs
    m = ref_loader.example_model("gp_regression")
    dim_y, dim_x = 12, 2
    data = {"x": rng.standard_normal((dim_y, dim_x)), "y_obs": np.zeros(dim_y)}
    dim_u = m.compute_dim_u(data)
    u_true = np.array([0.3, -0.2, 0.1, -1.0])[:dim_u]
    covar = np.asarray(m.covar_func(u_true, data))
    data["y_obs"] = np.linalg.cholesky(covar) @ rng.standard_normal(dim_y)
    system = build_system(mlift.GaussianProcessModelSystem, m, data, covar_func=m.covar_func)
    rec = {}

    def put(key, val):
        rec.setdefault(key, []).append(np.asarray(val, dtype=np.float64))

    for _ in range(4):
        u = u_true + 0.3 * rng.standard_normal(dim_u)
        chol_u = np.linalg.cholesky(np.asarray(m.covar_func(u, data)))
        q = np.concatenate([u, np.linalg.solve(chol_u, data["y_obs"])])
        up = u + 0.05 * rng.standard_normal(dim_u)
        chol_p = np.linalg.cholesky(np.asarray(m.covar_func(up, data)))
        qp = np.concatenate([up, np.linalg.solve(chol_p, data["y_obs"])
                             + 0.05 * rng.standard_normal(dim_y)])
        st, stp = state_of(q), state_of(qp)
        (dy_du, dy_dn), (dy_du_p, dy_dn_p) = system.jacob_constr_blocks(st), system.jacob_constr_blocks(stp)
        covar_u, dcovar_du = jax.vmap(
            partial(jax.jvp, lambda uu: m.covar_func(uu, data), (u,)), out_axes=(None, 1)
        )((np.identity(dim_u),))
        vq, wy = rng.standard_normal(q.size), rng.standard_normal(dim_y)
        put("q", q), put("q_prev", qp), put("vec_q", vq), put("vec_y", wy)
        put("covar", covar_u), put("dcovar_du", dcovar_du)
        put("c", system.constr(st)), put("dy_du", dy_du), put("dy_dn", dy_dn)
        put("dy_du_prev", dy_du_p), put("dy_dn_prev", dy_dn_p)
        put("chol_cap", system.gram_components(st)[0])
        put("logdet", system.log_det_sqrt_gram(state_of(q)))
        put("grad_logdet", system.grad_log_det_sqrt_gram(st))
        put("jv", system.lmult_by_jacob_constr(st, vq))
        put("jtw", system.rmult_by_jacob_constr(st, wy))
        put("pinv_w", system.lmult_by_pinv_jacob_constr(st, wy))
        put("ijp_w", system.lmult_by_inv_jacob_product(st, stp, wy))
        put("proj_v", system.project_onto_cotangent_space(vq.copy(), st))
    out.update({f"gp_{k}": np.array(v) for k, v in rec.items()})
    out["gp_x"], out["gp_y_obs"] = data["x"], data["y_obs"]
    np.savez(os.path.join(HERE, "ref_dense.npz"), **out)
    print("dense done")


if __name__ == "__main__":
    which = sys.argv[1:2] or ["linalg", "dense", "hier", "sv", "fn", "hh"]
    for w in which:
        {"hier": hier, "fn": fn, "hh": hh, "sv": sv, "linalg": linalg, "dense": dense}[w]()
