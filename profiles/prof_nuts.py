#!/usr/bin/env python
"""Dynamic multinomial HMC launches of the eight-schools workload at a FIXED step size, for
timing (CUDA events) and for `ncu -k regex:k_nuts` captures:

  python profiles/prof_nuts.py [--chains 65536] [--iters 20] [--launches 3] [--eps 0.66]

Prints one line per launch: ms, transitions, leapfrog steps, steps/s, mean depth."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import manifold_lifting_b200 as mlb  # noqa: E402
from manifold_lifting_b200.integrators import ConstrainedLeapfrogIntegrator  # noqa: E402
from manifold_lifting_b200.models import EightSchoolsModel  # noqa: E402
from manifold_lifting_b200.samplers import DynamicMultinomialHMC, StaticMetropolisHMC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=65536)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--warm", type=int, default=40)
ap.add_argument("--eps", type=float, default=0.66)
ap.add_argument("--static", type=int, default=0, help="n_step of the static sampler (0: dynamic)")
a = ap.parse_args()


class W:
    workload, n_school = "hier", 8


(y, s), q0, _ = bench.make_inputs(W, a.chains, 0)
model = EightSchoolsModel(y, s)
system = mlb.HierarchicalLatentVariableModelSystem(model, model.data, model.dim_u)
integ = ConstrainedLeapfrogIntegrator(
    system, a.eps, projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
    projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
if a.static:
    smp = StaticMetropolisHMC(system, integ, n_step=a.static, seed=7)
else:
    smp = DynamicMultinomialHMC(system, integ, max_tree_depth=10, seed=7)
state, _, _ = smp.sample_chains(0, a.warm, q0)
for _ in range(a.launches):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    state, _, st = smp.sample_chains(0, a.iters, state)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = float(st["n_step"].sum())
    print(f"{'static' if a.static else 'nuts'} chains {a.chains} iters {a.iters} ms {ms:.3f} "
          f"steps {steps:.4g} steps/s {steps / ms * 1e3:.4g} depth {float(st['tree_depth'].mean()):.3f} "
          f"accept {float(st['accept_stat'].mean()):.3f} iters/step "
          f"{float(st['projection_iters'].sum()) / max(steps, 1):.3f}", flush=True)
