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


if __name__ == "__main__":
    notebook_kat()
    autodiff_steps()
