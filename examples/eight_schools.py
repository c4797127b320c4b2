#!/usr/bin/env python
"""Eight-schools with manifold-lifted constrained HMC on a B200, end to end.

The batched counterpart of the reference's
`python mlift/example_models/eight_schools.py --algorithm chmc` (eight_schools.py:84-140,
driven by example_models/utils.py:run_experiment): model -> lifted system -> constrained
leapfrog integrator with the reference's Newton projection solver -> dynamic multinomial
HMC with dual-averaging step-size adaptation and the two-phase ToleranceAdapter -> traces,
`summary.json`, `final_chain_states.pkl` in --output-dir.  Instead of four chains in four
processes it runs --num-chain chains in one kernel launch per phase.

    python examples/eight_schools.py --num-chain 4096 --output-dir /tmp/eight_schools
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import manifold_lifting_b200 as mlb  # noqa: E402

Y_OBS = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])  # data/eight-schools-data.npz
SIGMA = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--num-chain", type=int, default=4096)
    ap.add_argument("--num-warm-up-iter", type=int, default=200)
    ap.add_argument("--num-main-iter", type=int, default=500)
    ap.add_argument("--prior-parametrization", default="unbounded", choices=("unbounded", "normal"))
    ap.add_argument("--step-size-adaptation-target", type=float, default=0.8)  # utils.py:82-87
    ap.add_argument("--max-tree-depth", type=int, default=10)  # utils.py:70-73
    ap.add_argument("--seed", type=int, default=202101)
    ap.add_argument("--output-dir", default=None)
    args = ap.parse_args(argv)

    rng = np.random.default_rng(args.seed)
    model = mlb.models.EightSchoolsModel(Y_OBS, SIGMA, args.prior_parametrization)
    system = mlb.HierarchicalLatentVariableModelSystem(model, model.data, model.dim_u)
    integrator = mlb.integrators.ConstrainedLeapfrogIntegrator(
        system, step_size=0.1,
        projection_solver=mlb.jitted_solve_projection_onto_manifold_newton,
        reverse_check_tol=2e-8,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    sampler = mlb.samplers.DynamicMultinomialHMC(
        system, integrator, max_tree_depth=args.max_tree_depth, seed=args.seed)
    adapters = [mlb.adapters.DualAveragingStepSizeAdapter(args.step_size_adaptation_target),
                mlb.ToleranceAdapter()]
    init = model.sample_initial_states(rng, args.num_chain)  # eight_schools.py:67-81

    t0 = time.time()
    final_state, traces, stats = sampler.sample_chains(
        args.num_warm_up_iter, args.num_main_iter, init, adapters=adapters,
        trace_dim=model.dim_u)
    import torch

    torch.cuda.synchronize()
    sampling_time = time.time() - t0

    u = traces["pos"]  # [n_chain, n_iter, 2]: (u0, u1) -> mu, tau (eight_schools.py:38-46)
    par = model.generate_params(u.reshape(-1, 2).cpu().numpy())
    named = {"μ": par["μ"].reshape(u.shape[:2]), "τ": par["τ"].reshape(u.shape[:2])}
    sub = min(args.num_chain, 512)  # diagnostics on a subset of chains keep this quick
    rows = {k: mlb.outputs.summarize_variable(v[:sub]) for k, v in named.items()}
    print(f"{args.num_chain} chains x {args.num_main_iter} draws in {sampling_time:.2f} s "
          f"(step size {float(integrator.step_size):.3f}, accept stat "
          f"{float(stats['accept_stat'].mean()):.3f}, mean tree depth "
          f"{float(stats['tree_depth'].mean()):.2f}, integrator errors "
          f"{float((stats['convergence_error'] + stats['non_reversible_step']).mean()):.2e})")
    for k, r in rows.items():
        print(f"  {k}: mean {r['mean']:.3f} sd {r['sd']:.3f} ess_bulk {r['ess_bulk']:.0f} "
              f"r_hat {r['r_hat']:.3f}")
    if args.output_dir:
        keep = {k: v[:sub] for k, v in named.items()}
        mlb.outputs.save_traces(args.output_dir, keep, {"accept_stat": stats["accept_stat"][:sub]})
        mlb.outputs.save_final_states(args.output_dir, final_state)
        mlb.outputs.compute_and_save_summary(
            args.output_dir, ["μ", "τ"], keep, total_sampling_time=sampling_time,
            final_integrator_step_size=float(integrator.step_size))
    return rows, stats


if __name__ == "__main__":
    main()
