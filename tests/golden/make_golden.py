#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ (run in the build container).

The reference (JAX 0.4.14 + Mici 0.2.1) cannot be imported here or on the GPU box, so the
fixtures come from two sources, both stated in each file's `source` field:

  * `notebook_euclidean_leapfrog.npz`: the ONLY numeric known-answer vector in the
    reference tree, the executed output of notebooks/Two-dimensional_example.ipynb
    (raw JSON lines 692-753): two Euclidean leapfrog steps on the toy posterior.
    Values are typed in from the notebook output, 8 printed digits.
  * `autodiff_steps_<model>.npz`: per-step outputs of the torch-fp64 AUTODIFF restatement
    (oracle/mlift_oracle: reference loops + autodiff Jacobians, i.e. the same
    construction as the reference's jax.jacfwd / jax.grad) for a handful of chains.  They
    pin the closed-form C++ oracle and the CUDA path against an independently derived
    implementation, NOT against the reference's own outputs ("parity unpinned").

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def notebook_kat():
    np.savez(
        os.path.join(HERE, "notebook_euclidean_leapfrog.npz"),
        source="notebooks/Two-dimensional_example.ipynb raw lines 692-753 (typed in)",
        sigma=0.1, y=1.0, step_size=0.05, coeff=2.0,
        pos=np.array([[-0.329414, -1.46108452], [-0.42719971, -1.06310332],
                      [-0.5247422, -0.65486249]]),
        mom=np.array([[-0.98537164, 0.25382918], [-1.95328196, 8.06222033],
                      [-2.12262924, 5.90596858]]),
    )  # fmt: skip


def autodiff_steps():
    from _cases import hier_case, hier_states, tangent_momentum, toy_case, toy_states
    from mlift_oracle import solvers as osolvers
    from mlift_oracle.mici import ChainState, ConstrainedLeapfrogIntegrator
    from mlift_oracle.mici import (mici_solve_projection_onto_manifold_newton,
                                   mici_solve_projection_onto_manifold_quasi_newton)

    solvers = {
        0: osolvers.solve_projection_onto_manifold_quasi_newton,
        1: osolvers.solve_projection_onto_manifold_newton,
        2: osolvers.solve_projection_onto_manifold_newton_with_line_search,
        3: mici_solve_projection_onto_manifold_newton,
        4: mici_solve_projection_onto_manifold_quasi_newton,
    }
    kw = dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50)
    for name, (omodel, osys, cm), states, sids in [
        ("hier_unbounded", hier_case("unbounded"), hier_states, (0, 1, 2)),
        ("hier_normal", hier_case("normal"), hier_states, (0, 1, 2)),
        ("toy", toy_case(0.1), toy_states, (3, 4)),
    ]:
        n, dt, n_step = 6, 0.15, 2
        q = states(omodel, n, 7)
        p = tangent_momentum(cm, q)
        out = {"q0": q, "p0": p, "dt": dt, "n_step": n_step, "solvers": np.array(sids)}
        for sid in sids:
            qs, ps, its, hs = [], [], [], []
            for i in range(n):
                integ = ConstrainedLeapfrogIntegrator(
                    osys, dt, projection_solver=solvers[sid], projection_solver_kwargs=kw)
                st = ChainState(q[i], p[i])
                it = []
                for _ in range(n_step):
                    st = integ.step(st)
                    it.append(integ.last_iters)
                qs.append(st.pos), ps.append(st.mom), its.append(it)
                hs.append(float(osys.h(st)))
            out[f"q_{sid}"], out[f"p_{sid}"] = np.array(qs), np.array(ps)
            out[f"iters_{sid}"], out[f"h_{sid}"] = np.array(its), np.array(hs)
        np.savez(os.path.join(HERE, f"autodiff_steps_{name}.npz"),
                 source="oracle/mlift_oracle autodiff restatement (see make_golden.py)",
                 **out)  # fmt: skip
        print(name, "done")


def fitzhugh_nagumo():
    """FN fixtures: (i) the arrays of the reference's data file the model needs
    (data/fitzhugh-nagumo-simulated-data.npz: t_seq, dt, y_mean_obs, n_obs - DATA, read at
    fitzhugh_nagumo.py:146-149), (ii) autodiff-oracle outputs at two states per
    parametrisation and one full constrained leapfrog step per solver."""
    import torch
    from mlift_oracle import solvers as osolvers
    from mlift_oracle.mici import ChainState, ConstrainedLeapfrogIntegrator
    from mlift_oracle.models import FitzHughNagumo
    from mlift_oracle.systems import IndependentAdditiveNoiseModelSystem

    d = np.load("/root/reference/data/fitzhugh-nagumo-simulated-data.npz")
    out = {"t_seq": d["t_seq"], "dt": d["dt"], "y_mean_obs": d["y_mean_obs"],
           "n_obs": d["n_obs"], "obs_noise_std": 0.1}  # fmt: skip
    y_obs = d["y_mean_obs"] + 0.1 * d["n_obs"]
    loc = np.array([0, 0, 0, -1, -3, -2, -1.0])
    u_true = np.log([3, 1, 3, 1 / 3, 1 / 15, 1 / 15, 0.1])
    rng = np.random.default_rng(0)
    for par in ("unbounded", "normal"):
        m = FitzHughNagumo(y_obs, d["t_seq"], d["dt"], par)
        osys = IndependentAdditiveNoiseModelSystem(
            m.generate_y, m.data, m.dim_u, m.extended_prior_neg_log_dens)
        qs, jac, dn, c, ld, gld, nld, gnld = ([] for _ in range(8))
        for _ in range(2):
            u = np.concatenate([u_true - (loc if par == "normal" else 0), [-1, 1]])
            q = m.lift(u + 0.05 * rng.standard_normal(9))
            q[9:] += 0.01 * rng.standard_normal(q.size - 9)  # off-manifold: c != 0
            st = ChainState(q)
            jb = osys.jacob_constr_blocks(st)
            qs.append(q), jac.append(np.asarray(jb[0])), dn.append(np.asarray(jb[1]))
            c.append(np.asarray(osys.constr(st)))
            ld.append(float(osys.log_det_sqrt_gram(st)))
            gld.append(np.asarray(osys.grad_log_det_sqrt_gram(st)))
            nld.append(float(osys.neg_log_dens(st)))
            gnld.append(np.asarray(osys.grad_neg_log_dens(st)))
        for k, v in dict(q=qs, dy_du=jac, dy_dn=dn, c=c, logdet=ld, grad_logdet=gld, nld=nld,
                         grad_nld=gnld).items():  # fmt: skip
            out[f"{par}_{k}"] = np.array(v)
        print("fn ops", par, "done")
        if par != "unbounded":
            continue
        q0 = m.lift(np.concatenate([u_true, [-1, 1]]) + 0.02 * rng.standard_normal(9))
        st0 = ChainState(q0)
        p0 = rng.standard_normal(q0.size)
        p0 = np.asarray(osys.project_onto_cotangent_space(torch.as_tensor(p0), st0))
        out["step_q0"], out["step_p0"], out["step_dt"] = q0, p0, 0.02
        for sid, fn in ((0, osolvers.solve_projection_onto_manifold_quasi_newton),
                        (1, osolvers.solve_projection_onto_manifold_newton)):  # fmt: skip
            integ = ConstrainedLeapfrogIntegrator(
                osys, 0.02, projection_solver=fn,
                projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8,
                                              max_iters=50))  # fmt: skip
            s1 = integ.step(ChainState(q0, p0))
            out[f"step_q_{sid}"], out[f"step_p_{sid}"] = np.asarray(s1.pos), np.asarray(s1.mom)
            out[f"step_iters_{sid}"] = np.array(integ.last_iters)
            out[f"step_h_{sid}"] = float(osys.h(s1))
            print("fn step solver", sid, integ.last_iters)
    np.savez(os.path.join(HERE, "fitzhugh_nagumo.npz"),
             source="data arrays: reference data/fitzhugh-nagumo-simulated-data.npz; "
                    "outputs: oracle/mlift_oracle autodiff restatement", **out)  # fmt: skip


def hodgkin_huxley():
    """HH fixtures: the constants / observations of the reference's data file
    (data/hodgkin-huxley-current-stimulus-data.npz, read at hh_current_stimulus.py:287-292)
    and autodiff-oracle outputs at one state in the spiking regime ("normal"
    parametrisation around the file's `u_obs`) plus one constrained leapfrog step per
    solver.  ~6 minutes of torch autodiff."""
    import torch
    from mlift_oracle import solvers as osolvers
    from mlift_oracle.mici import ChainState, ConstrainedLeapfrogIntegrator
    from mlift_oracle.models import HH_CONSTANT_KEYS, HodgkinHuxley
    from mlift_oracle.systems import IndependentAdditiveNoiseModelSystem

    d = dict(np.load("/root/reference/data/hodgkin-huxley-current-stimulus-data.npz"))
    out = {k: d[k] for k in HH_CONSTANT_KEYS}
    out.update(x_obs=d["x_obs"], n_obs=d["n_obs"], u_obs=d["u_obs"], obs_noise_std=1.0)
    y_obs = d["x_obs"] + 1.0 * d["n_obs"]
    rng = np.random.default_rng(0)
    par = "normal"
    m = HodgkinHuxley(y_obs, d, par)
    osys = IndependentAdditiveNoiseModelSystem(
        m.generate_y, m.data, m.dim_u, m.extended_prior_neg_log_dens)
    q = m.lift(d["u_obs"] + 0.02 * rng.standard_normal(7))
    q[7:] += 0.01 * rng.standard_normal(q.size - 7)
    st = ChainState(q)
    jb = osys.jacob_constr_blocks(st)
    out.update(q=q, dy_du=np.asarray(jb[0]), dy_dn=np.asarray(jb[1]),
               c=np.asarray(osys.constr(st)), logdet=float(osys.log_det_sqrt_gram(st)),
               grad_logdet=np.asarray(osys.grad_log_det_sqrt_gram(st)),
               nld=float(osys.neg_log_dens(st)),
               grad_nld=np.asarray(osys.grad_neg_log_dens(st)))  # fmt: skip
    print("hh ops done")
    q0 = m.lift(d["u_obs"] + 0.01 * rng.standard_normal(7))
    st0 = ChainState(q0)
    p0 = rng.standard_normal(q0.size)
    p0 = np.asarray(osys.project_onto_cotangent_space(torch.as_tensor(p0), st0))
    out["step_q0"], out["step_p0"], out["step_dt"] = q0, p0, 0.002
    for sid, fn in ((0, osolvers.solve_projection_onto_manifold_quasi_newton),
                    (1, osolvers.solve_projection_onto_manifold_newton)):  # fmt: skip
        integ = ConstrainedLeapfrogIntegrator(
            osys, 0.002, projection_solver=fn,
            projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
        s1 = integ.step(ChainState(q0, p0))
        out[f"step_q_{sid}"], out[f"step_p_{sid}"] = np.asarray(s1.pos), np.asarray(s1.mom)
        out[f"step_iters_{sid}"] = np.array(integ.last_iters)
        out[f"step_h_{sid}"] = float(osys.h(s1))
        print("hh step solver", sid, integ.last_iters)
    np.savez(os.path.join(HERE, "hodgkin_huxley.npz"),
             source="constants / x_obs / n_obs / u_obs: reference "
                    "data/hodgkin-huxley-current-stimulus-data.npz; outputs: "
                    "oracle/mlift_oracle autodiff restatement", **out)  # fmt: skip


if __name__ == "__main__":
    which = sys.argv[1:] or ["kat", "steps", "fn", "hh"]
    if "hh" in which:
        hodgkin_huxley()
    if "kat" in which:
        notebook_kat()
    if "steps" in which:
        autodiff_steps()
    if "fn" in which:
        fitzhugh_nagumo()
