#!/usr/bin/env python
"""Benchmark of the constrained-leapfrog hot path (BASELINE.json metric: constrained
leapfrog steps/s summed over all chains).

  python bench.py --gpus N --steps K --warmup W            # CUDA arm (this repo)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port)

A bench "step" is ONE `mlb_leapfrog_steps` call: `--n-step` constrained leapfrog steps
(A-B-A with projection and reversibility check) for every chain.
`value` = chains * n_step * K / (sum of the K CUDA-event intervals, max over ranks).
Prints ONE JSON line on rank 0.  Data are synthetic.

Default = BASELINE.json config 3 as written: the eight-schools-shaped model with 65,536
chains IN TOTAL, split into contiguous blocks of 65,536 / N chains per GPU ("scaling":
"strong").  The same line carries `weak` (65,536 chains per GPU, N > 1 only) and `legs`:
short runs of the other BASELINE configs - toy (65,536), FitzHugh-Nagumo (16,384) and
Hodgkin-Huxley (8,192 in total) - each with its own roofline / cpu_baseline / e2e.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SOLVERS = {"quasi-newton": 0, "newton": 1, "newton-line-search": 2}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mlb", choices=("mlb", "reference"))
    ap.add_argument("--workload", default="hier", choices=("hier", "toy", "fn", "hh", "sv"))
    ap.add_argument("--n-chain", type=int, default=0,
                    help="chains IN TOTAL under --scaling strong (default 65536; 16384 fn, "
                         "8192 hh), PER GPU under --scaling weak")  # fmt: skip
    ap.add_argument("--scaling", default="strong", choices=("strong", "weak"))
    ap.add_argument("--no-legs", action="store_true",
                    help="skip the extra legs (toy / fn / hh and the weak-scaling run)")
    ap.add_argument("--n-school", type=int, default=8)
    ap.add_argument("--n-step", type=int, default=0,
                    help="leapfrog steps per launch (default 32; 2 for fn)")  # fmt: skip
    ap.add_argument("--step-size", type=float, default=0.0,
                    help="default 0.1 (0.01 for fn)")  # fmt: skip
    ap.add_argument("--solver", default="newton", choices=tuple(SOLVERS))
    ap.add_argument("--cpu-sample-chains", type=int, default=0,
                    help="chains in the bounded CPU sample (0 = auto)")  # fmt: skip
    ap.add_argument("--single-kernel", action="store_true",
                    help="ODE workloads: one-thread-per-chain kernel instead of the "
                         "host-driven sub-phase kernels (MLB_FLAG_SINGLE_KERNEL)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ess", action="store_true", help="skip the ESS/s leg")
    ap.add_argument("--ess-sampler", default="both", choices=("both", "nuts", "static"),
                    help="samplers of the ESS/s leg: dynamic multinomial HMC (the reference's "
                         "default, utils.py:286-288; key `ess`) and / or static-trajectory "
                         "Metropolis HMC (key `ess_static`)")
    a = ap.parse_args()
    return with_workload_defaults(a, a.workload)


def with_workload_defaults(a, wl, **override):
    """Copy of the arguments with the per-workload defaults filled in (BASELINE configs)."""
    a = argparse.Namespace(**{**vars(a), "workload": wl, **override})
    a.n_chain = a.n_chain or {"fn": 16384, "hh": 8192, "sv": 16384}.get(wl, 65536)
    a.n_step = a.n_step or {"fn": 4, "hh": 4}.get(wl, 32)
    a.step_size = a.step_size or {"fn": 0.01, "hh": 0.002, "sv": 0.02}.get(wl, 0.1)
    return a


def local_chains(args, world):
    """(chains on each rank, chains in total): contiguous blocks of total / N under strong
    scaling, the full count per GPU under weak scaling."""
    if args.scaling == "weak":
        return args.n_chain, args.n_chain * world
    if args.n_chain % world:
        raise SystemExit(f"--n-chain {args.n_chain} is not divisible by {world} GPUs")
    return args.n_chain // world, args.n_chain


def workload_config(args, world):
    """The `config` object: what is run, from the arguments alone, so that the CUDA arm and
    the reference arm print the SAME dict (measured quantities go under `measured`)."""
    n_local, total = local_chains(args, world)
    return {
        "workload": workload_name(args), "chains_total": total, "chains_per_gpu": n_local,
        "n_step": args.n_step, "step_size": args.step_size, "solver": args.solver,
        "constraint_tol": 1e-9, "position_tol": 1e-8, "reverse_check_tol": 2e-8,
        "max_iters": 50,
        "l2": "state reset and L2 flushed (256 MiB memset) between timed steps, outside the "
              "CUDA-event intervals",
    }  # fmt: skip


# ------------------------------------------------------------------ workload (host)


def synthetic_schools(n_school, seed=202101):
    rng = np.random.default_rng(seed)
    sigma = rng.uniform(9.0, 18.0, n_school)
    y = 4.4 + 3.6 * rng.standard_normal(n_school) + sigma * rng.standard_normal(n_school)
    return y, sigma


FN_U_TRUE = np.concatenate([np.log([3.0, 1.0, 3.0, 1 / 3, 1 / 15, 1 / 15, 0.1]), [-1.0, 1.0]])


def synthetic_fn_data(seed=202101, n_obs=200, t_end=20.0, dt=0.05, noise_std=0.1):
    """simulate the FitzHugh-Nagumo ODE at the reference's true
    parameters (fitzhugh-nagumo-simulated-data.npz metadata) with host RK4 and add
    N(0, 0.1^2) observation noise."""
    a, b, g, d, e, z = np.exp(FN_U_TRUE[:6])
    t_seq = np.linspace(0.0, t_end, n_obs + 1)

    def f(x):
        return np.array([a * x[0] - b * x[0] ** 3 + g * x[1], -d * x[0] - e * x[1] + z])

    x, xs = FN_U_TRUE[7:9].copy(), []
    per = int(round((t_seq[1] - t_seq[0]) / dt))
    for _ in range(n_obs):
        for _ in range(per):
            k1 = f(x)
            k2 = f(x + dt * k1 / 2)
            k3 = f(x + dt * k2 / 2)
            k4 = f(x + dt * k3)
            x = x + dt * (k1 + 2 * k2 + 2 * k3 + k4) / 6
        xs.append(x[0])
    rng = np.random.default_rng(seed)
    return np.array(xs) + noise_std * rng.standard_normal(n_obs), t_seq, dt


def hh_constants():
    """Model constants of the Hodgkin-Huxley workload (rate constants, reversal potentials,
    stimulus timing) and the reference's nominal latent vector `u_obs` ("normal"
    parametrisation), from the committed fixture tests/golden/hodgkin_huxley.npz."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hodgkin_huxley.npz"))
    return {k: g[k] for k in g.files if g[k].ndim == 0 and k not in ("source",)}, g["u_obs"]


def synthetic_hh_data(forward, seed=202101, noise_std=1.0):
    """observations = membrane potential at the nominal parameters
    (computed by `forward(u)`: the implementation under test solves the ODE) + N(0, 1)."""
    consts, u_obs = hh_constants()
    x = forward(consts, u_obs)
    rng = np.random.default_rng(seed)
    return x + noise_std * rng.standard_normal(x.size), consts, u_obs


def make_inputs(args, n_chain, chain0):
    """On-manifold positions: hier - u ~ moderate prior draws, v ~ N(0, I),
    n = (y - x) / sigma; toy - theta ~ N(0, I) lifted.  Momenta ~ N(0, I) (projected on
    the device / by the oracle before use)."""
    rng = np.random.default_rng([202101, chain0])
    if args.workload == "hier":
        y, sigma = synthetic_schools(args.n_school)
        u = np.stack([rng.normal(0.0, 5.0, n_chain),
                      np.clip(np.log(np.abs(rng.standard_cauchy(n_chain) * 5.0)), -2.0, 3.0)], 1)  # fmt: skip
        v = rng.standard_normal((n_chain, args.n_school))
        x = u[:, :1] + np.exp(u[:, 1:2]) * v
        q = np.concatenate([u, v, (y[None] - x) / sigma[None]], 1)
        data = (y, sigma)
    elif args.workload == "fn":
        # u = truth + 0.1 N(0, I); the noise coordinates n = (y - x(u)) / sigma(u) are
        # filled in by the caller (needs an ODE solve: GPU `lift` / oracle `constr`)
        data = synthetic_fn_data()
        u = FN_U_TRUE[None] + 0.1 * rng.standard_normal((n_chain, 9))
        q = np.concatenate([u, np.zeros((n_chain, data[0].size))], 1)
    elif args.workload == "hh":
        # data and the noise coordinates need ODE solves: filled in by the caller
        consts, u_obs = hh_constants()
        data = None
        q = u_obs[None] + 0.05 * rng.standard_normal((n_chain, 7))
        q = np.concatenate([q, np.zeros((n_chain, 1001))], 1)
    else:
        sig = 0.1
        th = rng.standard_normal((n_chain, 2))
        f = th[:, 1] ** 2 + 2.0 * th[:, 0] ** 2 * (th[:, 0] ** 2 - 0.5)
        q = np.concatenate([th, ((1.0 - f) / sig)[:, None]], 1)
        data = (sig, 1.0, 2.0)
    p = rng.standard_normal(q.shape)
    return data, q, p


def workload_name(args):
    if args.workload == "hier":
        return (f"eight-schools-shaped hierarchical model (J={args.n_school}, dim_q="
                f"{2 + 2 * args.n_school}), synthetic data")  # fmt: skip
    if args.workload == "fn":
        return (f"FitzHugh-Nagumo ODE model (RK4 dt=0.05, 200 obs, dim_q=209), "
                f"synthetic observations")  # fmt: skip
    if args.workload == "hh":
        return (f"Hodgkin-Huxley current-stimulus model (1100 exp-Euler steps, 1001 obs, "
                f"dim_q=1008, normal parametrisation), synthetic observations")  # fmt: skip
    return "toy 2-D model (sigma=0.1, dim_q=3)"


# ------------------------------------------------------------------ CPU (oracle) leg


def cpu_oracle_rate(args, target_seconds=12.0, sample_chains=0):
    """Times the C++ oracle port on all host cores on a bounded sample of the workload.
    Returns (steps_per_s, cores, sample_description, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as co

    cores = co.use_all_cores()
    n = sample_chains or {"fn": 16 * cores, "hh": 4 * cores}.get(
        args.workload, 2048 * max(1, cores // 4))
    data, q, p = make_inputs(args, n, 10**6)
    if args.workload == "hier":
        cm = co.OracleModel.hier(*data)
    elif args.workload == "fn":
        cm = co.OracleModel.fn(*data)
        q[:, 9:] = -cm.constr(q) / np.exp(q[:, 6:7])  # lift onto the manifold
    elif args.workload == "hh":
        def forward(consts, u):
            m0 = co.OracleModel.hh(np.zeros(1001), consts, "normal")
            return m0.constr(np.concatenate([u, np.zeros(1001)])[None])[0]

        y_obs, consts, _ = synthetic_hh_data(forward)
        cm = co.OracleModel.hh(y_obs, consts, "normal")
        q[:, 7:] = -cm.constr(q) / np.exp(q[:, 6:7])
    else:
        cm = co.OracleModel.toy(*data)
    jac, _ = cm.jacob_constr_blocks(q)
    p = p - cm.normal_space_component(jac, cm.decompose_gram(jac), p)
    opts = co.solver_opts(SOLVERS[args.solver])
    nw = min(n, cores if args.workload in ("fn", "hh") else 64 * cores)
    cm.leapfrog_steps(q[:nw], p[:nw], args.step_size, 1, opts)  # warm
    t0 = time.perf_counter()
    r = cm.leapfrog_steps(q, p, args.step_size, args.n_step, opts)
    dt = time.perf_counter() - t0
    done = int(r["n_done"].sum())
    reps = 1
    if dt < target_seconds / 4:  # enlarge sample towards the target duration
        reps = int(min(50, target_seconds / max(dt, 1e-3)))
        t0 = time.perf_counter()
        for _ in range(reps):
            r = cm.leapfrog_steps(q, p, args.step_size, args.n_step, opts)
        dt = time.perf_counter() - t0
        done = int(r["n_done"].sum()) * reps
    desc = (f"{n} chains x {args.n_step} steps x {reps} repeats of the same workload, "
            f"C++ oracle port, OpenMP over chains")  # fmt: skip
    return done / dt, cores, desc, dt


def run_reference(args, emit):
    """`--impl reference`: the reference's CPU implementation of the path - here the C++
    oracle port on all host cores (JAX / Mici are not installable offline) - on the CUDA
    arm's config, metric and unit; each bench step is a bounded sample of the workload.
    Under torchrun rank 0 alone runs it; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.workload == "sv":
        return run_reference_sv(args, emit)
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        r, cores, desc, dt = cpu_oracle_rate(args, target_seconds=1.0,
                                             sample_chains=args.cpu_sample_chains)  # fmt: skip
        if i >= args.warmup:
            rates.append(r), secs.append(dt)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "constrained leapfrog steps/s (all chains)",
        "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, max(world, args.gpus)),
        "note": "reference arm = CPU oracle port on host cores (JAX/Mici are not installable "
                "offline); each bench step is a bounded sample of the config's workload",
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    emit(line)


# ------------------------------------------------------------------------ CUDA leg


# ------------------------------------------------------------------------ ESS/s leg

ESS_WARM, ESS_MAIN, ESS_N_STEP, ESS_SUBSET, ESS_SEED = 150, 200, 8, 8192, 202101
ESS_MAX_DEPTH = 10  # max_tree_depth default of the reference (utils.py:70-73)


def ess_leg(args, mlb, model, system, q0, rank, world, n, with_cpu, nuts=True,
            oracle_model=None):
    """Second half of BASELINE.json's metric: bulk-ESS per second of the latent parameters
    under constrained HMC - by default the reference's sampler, dynamic multinomial HMC
    (utils.py:286-288; `--ess-sampler static` for static-trajectory Metropolis HMC with
    n_step=8) - with dual-averaging warm-up to accept statistic 0.8, GPU sampler vs the CPU
    oracle sampler.

    The CPU leg continues the SAME chains (Philox streams are keyed by seed / global chain
    id / iteration) from the GPU's post-warm-up states, so both legs have identical ESS
    per chain-iteration and the ESS/s ratio isolates throughput; the max |difference| of
    the two traces is reported as a sampler-level parity figure.  ESS is estimated on
    the first ESS_SUBSET chains of each rank and scaled to the chain count."""
    import torch

    ode = args.workload in ("fn", "hh")
    ESS_WARM, ESS_MAIN, ESS_N_STEP = (30, 40, 4) if ode else (150, 200, 8)
    if ode and nuts:
        ESS_WARM, ESS_MAIN = 15, 20

    from manifold_lifting_b200 import parallel
    from manifold_lifting_b200.adapters import DualAveragingStepSizeAdapter
    from manifold_lifting_b200.integrators import ConstrainedLeapfrogIntegrator
    from manifold_lifting_b200.samplers import DynamicMultinomialHMC, StaticMetropolisHMC
    from manifold_lifting_b200.solvers import SOLVER_IDS

    solver_fn = {v: k for k, v in SOLVER_IDS.items()}[SOLVERS[args.solver]]
    integ = ConstrainedLeapfrogIntegrator(
        system, args.step_size, projection_solver=solver_fn,
        projection_solver_kwargs=dict(constraint_tol=1e-9, position_tol=1e-8, max_iters=50))
    if nuts:
        sampler = DynamicMultinomialHMC(system, integ, max_tree_depth=ESS_MAX_DEPTH, seed=ESS_SEED)
    else:
        sampler = StaticMetropolisHMC(system, integ, n_step=ESS_N_STEP, seed=ESS_SEED)
    state, _, _ = sampler.sample_chains(ESS_WARM, 0, q0, chain_offset=rank * n,
                                        adapters=[DualAveragingStepSizeAdapter(0.8, 0.05)])  # fmt: skip
    eps = float(integ.step_size)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, traces, stats = sampler.sample_chains(0, ESS_MAIN, state, chain_offset=rank * n,
                                             trace_dim=model.dim_u)  # fmt: skip
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    t = torch.tensor([secs], dtype=torch.float64, device="cuda")
    parallel.all_reduce_max_(t)
    secs = float(t[0])
    u_tr = traces["pos"]  # [n, ESS_MAIN, dim_u]
    sub = min(ESS_SUBSET, n)
    names = ("u0", "u1") if model.dim_u == 2 else tuple(f"u{k}" for k in range(model.dim_u))
    scale = n * world / (sub * world)
    ess = {nm: parallel.ess(u_tr[:sub, :, k].contiguous(), "bulk") * scale
           for k, nm in enumerate(names)}  # fmt: skip
    rhat = {nm: parallel.rhat(u_tr[:sub, :, k].contiguous(), "rank") for k, nm in enumerate(names)}
    acc = torch.stack([stats["accept_stat"].sum(), stats["convergence_error"].sum()
                       + stats["non_reversible_step"].sum()])  # fmt: skip
    parallel.all_reduce_sum_(acc)
    out = {
        "sampler": (f"dynamic multinomial HMC (NUTS), max_tree_depth={ESS_MAX_DEPTH}" if nuts
                    else f"static HMC, n_step={ESS_N_STEP}")
                   + f", {ESS_WARM} warm-up + {ESS_MAIN} main iterations, dual-averaging "
                     f"target 0.8",
        "mean_tree_depth": float(stats["tree_depth"].mean()) if nuts else None,
        "mean_steps_per_transition": float(stats["n_step"].mean()) / ESS_MAIN,
        "step_size": eps, "accept_stat": float(acc[0]) / (n * world),
        "integrator_error_rate": float(acc[1]) / (n * world),
        "seconds_main": secs, "chains": n * world,
        "ess_bulk": ess, "r_hat": rhat, "ess_per_s": {k: v / secs for k, v in ess.items()},
        "ess_estimated_on_chains": sub * world,
    }  # fmt: skip
    if with_cpu and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import c_oracle as co

        cores = co.use_all_cores()
        n_cpu = min(sub, (2 if ode else 64) * cores)
        if oracle_model is not None:
            cm = oracle_model(co)
        elif args.workload == "hier":
            cm = co.OracleModel.hier(model.data["y_obs"], model.data["σ"])
        else:
            cm = co.OracleModel.toy(model.sigma, model.y, model.coeff)
        qc = state.pos[:n_cpu].cpu().numpy().copy()
        opts = co.solver_opts(SOLVERS[args.solver])
        tr = np.empty((n_cpu, ESS_MAIN, model.dim_u))
        t0 = time.perf_counter()
        for it in range(ESS_MAIN):
            if nuts:
                r = cm.nuts_transition(qc, eps, opts, max_depth=ESS_MAX_DEPTH, seed=ESS_SEED,
                                       chain0=0, iteration=ESS_WARM + it)  # fmt: skip
            else:
                r = cm.hmc_transition(qc, eps, ESS_N_STEP, opts, seed=ESS_SEED, chain0=0,
                                      iteration=ESS_WARM + it)  # fmt: skip
            qc = r["q"]
            tr[:, it] = qc[:, : model.dim_u]
        csecs = time.perf_counter() - t0
        tr_t = torch.as_tensor(tr)
        cess = {nm: float(_single_process_ess(tr_t[:, :, k])) for k, nm in enumerate(names)}
        gpu_sub = u_tr[:n_cpu].cpu().numpy()
        same = np.abs(gpu_sub - tr).max(axis=(1, 2)) < 1e-6
        out["cpu"] = {
            "chains": n_cpu, "cores": cores, "seconds_main": csecs, "ess_bulk": cess,
            "ess_per_s": {k: v / csecs for k, v in cess.items()},
            "same_chains_as_gpu_fraction": float(same.mean()),
            # chains whose CPU and GPU traces stay within 1e-6 for all iterations (a
            # borderline accept decision or iteration-count tie makes a chain diverge;
            # on strongly curved manifolds such as the toy model most eventually do)
            "max_abs_trace_diff_on_same_chains": (float(np.abs(gpu_sub - tr)[same].max())
                                                  if same.any() else None),
        }
        out["ess_per_s_ratio_gpu_over_cpu"] = {
            k: out["ess_per_s"][k] / out["cpu"]["ess_per_s"][k] for k in cess}
    return out


def _single_process_ess(x):
    """parallel.ess without collectives (the CPU leg runs on rank 0 only)."""
    from manifold_lifting_b200 import parallel as par

    xs = par._split(x.to(dtype=__import__("torch").float64))
    n_draw = xs.shape[1]
    flat = xs.reshape(-1)
    import torch

    order = torch.argsort(flat, stable=True)
    ranks = torch.empty_like(flat)
    ranks[order] = torch.arange(1, flat.numel() + 1, dtype=flat.dtype)
    z = (ranks - 0.375) / (flat.numel() + 0.25)
    z = (2.0**0.5) * torch.erfinv(2.0 * z - 1.0)
    stats, ac = par._sufficient(z.reshape(xs.shape))
    return par._ess_from(stats, ac, n_draw)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML polled from a
    host thread every 2 ms; the `nvidia-smi --query-gpu=clocks.sm,...` fields of
    B200_PROFILING.md read through the library instead of a subprocess, so that a timed
    region of tens of milliseconds still gets samples)."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80))  # fmt: skip

    def __init__(self, index, period=0.002):
        # the ODE workloads are driven from the host (hundreds of launches and a few
        # stream synchronisations per step) and NVML queries contend with those calls
        # inside the driver: they are polled every 25 ms instead of every 2 ms
        self.period = period
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.replace(",", "").isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), int(rs)))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self.stop_flag = True
        self.thread.join(1.0)
        sm = [r[0] for r in self.rows]
        reasons = sorted({name for _, bits in self.rows for name, mask in self.REASONS if bits & mask})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max,
                "reasons": reasons, "samples": len(sm)}  # fmt: skip


def run_mlb(args, emit):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the mlb arm)")
    torch.cuda.set_device(local)
    bound_cpus = 0
    if world > 1:  # one process per GPU: stay on the cores (and NUMA node) next to it
        from manifold_lifting_b200 import parallel as _par

        bound_cpus = _par.bind_to_gpu_cpus(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = {"rank": rank, "world": world, "local": local, "bound_cpus": bound_cpus}

    line = measure(args, ctx, full=True)
    if not args.no_legs:
        # the other BASELINE configs, a few steps each, same keys (no ESS legs)
        k_leg = max(2, min(args.steps, 5))
        if world > 1 and args.scaling == "strong":
            weak = measure(with_workload_defaults(args, args.workload, scaling="weak",
                                                  n_chain=local_chains(args, 1)[0], steps=k_leg),
                           ctx, full=False)  # fmt: skip
            if rank == 0:
                line["weak"] = {k: weak[k] for k in ("value", "unit", "ms_per_step", "e2e",
                                                     "config", "roofline")}  # fmt: skip
        legs = {}
        for wl in ("toy", "fn", "hh"):
            if wl == args.workload:
                continue
            a = with_workload_defaults(args, wl, n_chain=0, n_step=0, step_size=0.0,
                                       steps=k_leg, scaling="strong")
            leg = measure(a, ctx, full=False)
            if rank == 0:
                legs[wl] = leg
        if rank == 0:
            line["legs"] = legs
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure(args, ctx, full):
    """One workload on this process's GPU (all ranks call it together): device-timed value,
    the one-step-per-launch HBM view, end to end through the host-buffer entry point, the
    CPU oracle on rank 0; `full` adds the ESS legs.  Returns the JSON object on rank 0."""
    import torch
    import torch.distributed as dist

    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib, flops

    rank, world, local = ctx["rank"], ctx["world"], ctx["local"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    n, total = local_chains(args, world)
    n_step, K, W = args.n_step, args.steps, max(args.warmup, 3)
    # (MLB_BENCH_RANK: draw the inputs of another rank's share on a single GPU - diagnostics)
    data, q_host, p_host = make_inputs(args, n, int(os.environ.get("MLB_BENCH_RANK", rank)) * n)
    y_obs = consts = None
    if args.workload == "hier":
        model = mlb.models.EightSchoolsModel(*data)
        system = mlb.HierarchicalLatentVariableModelSystem(model, model.data, model.dim_u)
    elif args.workload == "fn":
        model = mlb.models.FitzHughNagumoModel(*data)
        system = mlb.IndependentAdditiveNoiseModelSystem(model, model.data, model.dim_u)
        q_host = model.lift(q_host[:, :9])
    elif args.workload == "hh":
        def forward(consts, u):
            m0 = mlb.models.HodgkinHuxleyModel(np.zeros(1001), consts, "normal")
            return -m0.lift(u[None])[0, 7:].cpu().numpy() * np.exp(u[6])

        y_obs, consts, _ = synthetic_hh_data(forward)
        model = mlb.models.HodgkinHuxleyModel(y_obs, consts, "normal")
        system = mlb.IndependentAdditiveNoiseModelSystem(model, model.data, model.dim_u)
        q_host = model.lift(q_host[:, :7])
    else:
        model = mlb.models.Toy2DModel(*data)
        system = mlb.IndependentAdditiveNoiseModelSystem(model, model.data, model.dim_u)
    Q = model.dim_q
    L = _lib.lib()
    solver = SOLVERS[args.solver]
    opts = _lib.SolverOpts(solver, 50, 10, 0, 1e-9, 1e-8, 1e10, 2e-8, 0,
                           _lib.FLAG_SINGLE_KERNEL if args.single_kernel else 0)

    q0 = _lib.to_soa(q_host)
    st0 = mlb.states.ChainState(q0)
    p0 = system.project_onto_cotangent_space(_lib.to_soa(p_host), st0)
    q, p = _lib.soa_clone(q0), _lib.soa_clone(p0)
    status, n_done, itf, itr = (torch.empty(n, dtype=torch.int32, device="cuda")
                                for _ in range(4))  # fmt: skip
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def launch(nstep):
        pq, ld = _lib.p_soa(q, Q)
        _lib.check(L.mlb_leapfrog_steps(
            model.handle, C.c_int64(n), C.c_int64(ld), pq, _lib.p_soa(p, Q)[0],
            C.c_double(args.step_size), None, C.c_int32(nstep), C.byref(opts),
            _lib.p_vec(status, torch.int32), _lib.p_vec(n_done, torch.int32),
            _lib.p_vec(itf, torch.int32), _lib.p_vec(itr, torch.int32), None, None,
            C.c_void_p(stream.cuda_stream)))  # fmt: skip

    def reset():
        q.copy_(q0), p.copy_(p0)
        flush.zero_()

    def timed(nstep, iters):
        evs = []
        for _ in range(iters):
            reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launch(nstep)
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    for _ in range(W):
        reset(), launch(n_step)
    barrier()
    clocks = ClockSampler(local, 0.025 if args.workload in ("fn", "hh") else 0.002)
    clocks.start()
    launches0 = L.mlb_launch_count()
    barrier()
    ms = timed(n_step, K)
    barrier()
    launches = L.mlb_launch_count() - launches0
    clk = clocks.stop()
    total_ms = float(sum(ms))
    done_local = float(n_done.sum())  # chain-steps completed in the LAST timed step
    mean_iters = float((itf.sum() + itr.sum()) / max(done_local, 1.0))
    failed = float((status != 0).double().mean())

    # --- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timing)
    qh = torch.empty(n, Q, dtype=torch.float64).pin_memory()
    ph = torch.empty(n, Q, dtype=torch.float64).pin_memory()
    q_src, p_src = q0.cpu().contiguous(), p0.cpu().contiguous()
    sth = torch.empty(n, dtype=torch.int32).pin_memory()
    ndh = torch.empty(n, dtype=torch.int32).pin_memory()
    e2e_iters = max(3, min(K, 10))
    e2e_s = []
    for i in range(2 + e2e_iters):
        qh.copy_(q_src), ph.copy_(p_src)
        barrier()
        t0 = time.perf_counter()
        _lib.check(L.mlb_leapfrog_steps_host(
            model.handle, C.c_int64(n), C.c_void_p(qh.data_ptr()), C.c_void_p(ph.data_ptr()),
            C.c_double(args.step_size), C.c_int32(n_step), C.byref(opts),
            C.c_void_p(sth.data_ptr()), C.c_void_p(ndh.data_ptr()), None, None, None, None))  # fmt: skip
        torch.cuda.synchronize()
        if i >= 2:
            e2e_s.append(time.perf_counter() - t0)
    e2e_done = float(ndh.sum())
    e2e_time = float(np.mean(e2e_s))

    # --- HBM view: one leapfrog step per launch, L2 flushed, same chains
    ms1 = timed(1, max(5, min(K, 20)))
    done1 = float(n_done.sum())
    fp64_peak = L.mlb_measure_fp64_tflops(3) if rank == 0 else 0.0

    red = torch.tensor([total_ms, e2e_time, float(np.mean(ms1))], dtype=torch.float64,
                       device="cuda")  # fmt: skip
    tot = torch.tensor([done_local * K, e2e_done, done1], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    line = None
    if rank == 0:
        total_ms, e2e_time, ms1_mean = (float(x) for x in red)
        steps_all, e2e_all, done1_all = (float(x) for x in tot)
        value = steps_all / (total_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        flops_step = flops.step_flops(model, solver, mean_iters)
        if args.workload in ("hier", "toy"):
            kernel_name = "k_leapfrog_steps (fused trajectory, one launch)"
        elif args.single_kernel:
            kernel_name = "ks_leapfrog_steps (fused trajectory, one launch)"
        else:
            kernel_name = ("mlb_leapfrog_steps call = host-driven sub-phase kernels; kp_jacob and "
                           "kp_second_order (sensitivity sweeps) dominate, see profiles/")
        # DRAM bytes of the dominant kernel from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                args.workload, {})
        except (OSError, ValueError):
            traffic = {}
        ach_tf = flops_step * done_local / (float(np.mean(ms)) * 1e-3) / 1e12
        bytes_step = flops.step_bytes(model)
        ach_gbs = bytes_step * (done1_all / world) / (ms1_mean * 1e-3) / 1e9
        line = {
            "metric": "constrained leapfrog steps/s (all chains)", "value": value,
            "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "measured": {
                "mean_projection_iters_per_step": mean_iters,
                "failed_chain_fraction": failed,
                "cpus_bound_per_rank": ctx["bound_cpus"] or None,
                "timing": "sum of per-step CUDA-event intervals on the launch stream, max "
                          "over ranks",
            },
            "roofline": {
                "bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp64_peak if fp64_peak > 0 else None,
                "traffic": traffic.get("bytes_per_launch"),
                "traffic_source": traffic.get("source"),
                "kernel": kernel_name,
                "algorithmic_flops_per_chain_step": flops_step,
                "chains_this_gpu": n,
                "peak_source": "DFMA-loop microbenchmark run live in this process "
                               "(MEASURED_PEAKS.json has no fp64 entry)",
            },
            "roofline_hbm": {
                "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach_gbs / hbm_peak, "traffic": None,
                "kernel": kernel_name + ", one step per launch, L2 flushed",
                "algorithmic_bytes_per_chain_step": bytes_step,
                "ms_per_launch": ms1_mean, "peak_source": hbm_src,
            },
            "e2e": {"value": e2e_all / e2e_time, "unit": "steps/s",
                    "h2d_bytes_per_step": 2 * n * Q * 8,
                    "d2h_bytes_per_step": 2 * n * Q * 8 + 2 * n * 4,
                    "h2d_gbs_per_rank": 2 * n * Q * 8 / e2e_time / 1e9,
                    "d2h_gbs_per_rank": (2 * n * Q * 8 + 2 * n * 4) / e2e_time / 1e9,
                    "api": "mlb_leapfrog_steps_host (pinned host [n, Q] buffers in/out)"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }  # fmt: skip
        if not args.no_cpu_baseline:
            r, cores, desc, secs = cpu_oracle_rate(args, 12.0 if full else 4.0,
                                                   args.cpu_sample_chains)  # fmt: skip
            line["cpu_baseline"] = {"value": r, "unit": "steps/s", "cores": cores,
                                    "kind": "port", "sample": desc, "seconds": secs}  # fmt: skip
    ess = {}
    if full and not args.no_ess and args.workload in ("hier", "toy"):
        for key, nuts in (("ess", True), ("ess_static", False)):
            if args.ess_sampler in ("both", "nuts" if nuts else "static"):
                ess[key] = ess_leg(args, mlb, model, system, q0, rank, world, n,
                                   with_cpu=not args.no_cpu_baseline and world == 1,
                                   nuts=nuts)  # fmt: skip
    elif full and not args.no_ess and args.workload in ("fn", "hh"):
        # static sampler by default; the host-driven dynamic sampler (a transition costs as
        # many lock-step rounds as the largest tree of the batch) with --ess-sampler nuts
        for key, nuts in (("ess", True), ("ess_static", False)):
            if args.ess_sampler == ("nuts" if nuts else "static") or (
                    args.ess_sampler == "both" and not nuts):
                ess[key] = ess_leg(
                    args, mlb, model, system, q0, rank, world, n,
                    with_cpu=not args.no_cpu_baseline and world == 1, nuts=nuts,
                    oracle_model=(lambda co: co.OracleModel.fn(*data)) if args.workload == "fn"
                    else (lambda co: co.OracleModel.hh(y_obs, consts, "normal")))  # fmt: skip
    if rank == 0:
        line.update({k: v for k, v in ess.items() if v is not None})
    del q, p, q0, p0, flush, qh, ph
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------- state-space workload (extra, `--workload sv`)


def run_reference_sv(args, emit):
    """CPU arm of `--workload sv`: the NumPy restatement (one core) on a bounded sample of
    chains started on the manifold near the generating parameters."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from mlift_oracle import state_space as oss

    y = sv_synthetic_observations()
    solver = {"quasi-newton": "quasi_newton", "newton": "newton"}[args.solver]
    rng = np.random.default_rng(202101)
    m = args.cpu_sample_chains or 4
    qs, _ = oss.sv_on_manifold_states(rng, m, SV_DIM_Y, y=y)
    ps = []
    for q in qs:
        jac, _ = oss.sv_jacob_constr_split_blocks(*oss.sv_split(q, SV_DIM_Y), y)
        p = rng.normal(size=q.size)
        ps.append(p - oss.rmult_by_jacob_constr(*jac, oss.lmult_by_inv_gram(
            *jac, *oss.decompose_gram(*jac), oss.lmult_by_jacob_constr(*jac, p))))
    secs = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for q, p in zip(qs, ps):
            oss.sv_leapfrog_step(q, p, args.step_size, y, solver=solver)
        if i >= args.warmup:
            secs.append(time.perf_counter() - t0)
    value = m / float(np.mean(secs))
    emit({
        "impl": "reference", "metric": "constrained leapfrog steps/s (all chains)",
        "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"stochastic-volatility state-space system (dim_y={SV_DIM_Y}), "
                               "extra workload outside the BASELINE configs",
                   "step_size": args.step_size, "solver": args.solver},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": 1, "kind": "port",
                         "sample": f"{m} chains x 1 constrained leapfrog step per bench step, NumPy "
                                   "restatement, one core"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })  # fmt: skip


SV_DIM_Y = 1000


def sv_synthetic_observations(dim_y=SV_DIM_Y, seed=202101):
    """Stochastic-volatility observations (stochastic_volatility.py:29-38) at mu = -0.5,
    sigma = 0.3, phi = 0.76; observation noise kept away from zero (|n| >= 0.2)."""
    rng = np.random.default_rng(seed)
    mu, sigma, phi = -0.5, 0.3, -1.0 + 2.0 / (1.0 + np.exp(-2.0))
    v, z = rng.normal(size=dim_y), rng.normal(size=dim_y)
    noise = np.where(z < 0, -1.0, 1.0) * (0.2 + np.abs(z))
    x = np.empty(dim_y)
    x[0] = mu + sigma / np.sqrt(1 - phi**2) * v[0]
    for t in range(1, dim_y):
        x[t] = mu + phi * (x[t - 1] - mu) + sigma * v[t]
    return np.exp(x / 2) * noise


def run_sv(args, emit):
    """Extra workload outside the BASELINE configs: one constrained
    leapfrog step (half kick with dh1_dpos incl. the gradient of log_det_sqrt_gram +
    cotangent projection, position step with projection onto the manifold, momentum
    re-projection, reversibility check, second half kick) for the stochastic-volatility
    state-space system, dim_y = 1,000, every chain."""
    import torch
    import torch.distributed as dist

    import manifold_lifting_b200 as mlb
    from manifold_lifting_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the mlb arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ss = mlb.state_space
    n, dim_y, dt = args.n_chain, SV_DIM_Y, args.step_size
    solver = {"quasi-newton": "quasi_newton", "newton": "newton"}.get(args.solver)
    if solver is None:
        raise SystemExit("--workload sv supports --solver newton | quasi-newton")
    y = sv_synthetic_observations()
    model = ss.StochasticVolatilityModel(y)
    g = torch.Generator(device="cuda").manual_seed(202101 + rank)
    q = _lib.soa_empty(n, model.dim_q).normal_(generator=g)
    q[:, 0] = -0.5 + 0.2 * q[:, 0]
    q[:, 1] = float(np.log(0.3)) + 0.2 * q[:, 1]
    q[:, 2] = 2.0 + 0.2 * q[:, 2]
    sigma, phi = q[:, 1].exp(), -1.0 + 2.0 * torch.sigmoid(q[:, 2])
    yt = torch.as_tensor(y).cuda()
    x = q[:, 0] + sigma / (1 - phi**2).sqrt() * q[:, 3]
    for t in range(dim_y):  # lift onto the manifold: n_t = y_t / exp(x_t / 2)
        if t > 0:
            x = q[:, 0] + phi * (x - q[:, 0]) + sigma * q[:, 3 + t]
        q[:, 3 + dim_y + t] = yt[t] / (x / 2).exp()
    J, _ = model.jacob_constr_blocks(q)
    gram = ss.decompose_gram(J)
    p = ss.project_onto_cotangent_space(
        _lib.soa_empty(n, model.dim_q).normal_(generator=g), J, gram)
    del J, gram

    def step():
        return model.leapfrog_step(q, p, dt, solver=solver)

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    clocks = ClockSampler(local, 0.025)
    clocks.start()
    l0 = mlb._lib.lib().mlb_launch_count()
    ms = 0.0
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    launches = int(mlb._lib.lib().mlb_launch_count() - l0)
    clk = clocks.stop()
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ok_frac = float((out["status"] == _lib.ST_OK).double().mean())
    value = n * world * args.steps / (ms * 1e-3)

    # end to end: pinned host [n, Q] arrays in, positions / momenta / status out
    # (host arrays in the component-major layout too, so each copy is one contiguous transfer)
    def pinned():
        return torch.empty(model.dim_q, n, dtype=torch.float64, pin_memory=True).t()

    qh, ph, qo, po = pinned(), pinned(), pinned(), pinned()
    qh.copy_(q), ph.copy_(p)
    e2e_iters = max(3, args.steps // 3)

    def e2e_step():
        o = model.leapfrog_step(qh.cuda(non_blocking=True), ph.cuda(non_blocking=True), dt,
                                solver=solver)
        qo.copy_(o["pos"], non_blocking=True), po.copy_(o["mom"], non_blocking=True)
        return o["status"].cpu()

    e2e_step()  # untimed: first-use allocations of the staging tensors
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_iters):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_iters / float(te.item())

    # roofline of the dominant kernel, timed alone on this stream
    Jn, _ = model.jacob_constr_blocks(out["pos"])
    c = model.constr(q + dt * p)
    if solver == "newton":
        kern = lambda: ss.lmult_by_inv_jacob_product(Jn, None, None, None, Jn, None, None, None, c)  # noqa: E731
        kname, rows = "k_ssm_inv_jacob_product<3>", 2 * 3 + 8
    else:
        gn = ss.decompose_gram(Jn)
        kern = lambda: ss.lmult_by_inv_gram(Jn, None, None, None, *gn, c)  # noqa: E731
        kname, rows = "k_ssm_inv_gram<3>", 3 + 4
    kern()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        kern()
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / 10
    alg_bytes = 8.0 * n * rows * dim_y
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        hbm_peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    ach = alg_bytes / (kms * 1e-3) / 1e9

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from mlift_oracle import state_space as oss

        m = args.cpu_sample_chains or 12
        qs, ps = q[:m].cpu().numpy(), p[:m].cpu().numpy()
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            for i in range(m):
                oss.sv_leapfrog_step(qs[i], ps[i], dt, y, solver=solver)
        sec = time.perf_counter() - t0
        cpu = {"value": m / sec, "unit": "steps/s", "cores": 1, "kind": "port",
               "sample": f"{m} chains x 1 constrained leapfrog step, NumPy restatement "
                         "(oracle/mlift_oracle/state_space.py), one core", "seconds": sec}  # fmt: skip
    if rank == 0:
        emit({
            "metric": "constrained leapfrog steps/s (all chains)", "value": value,
            "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"stochastic-volatility state-space system (dim_u=3, dim_y={dim_y}, "
                                   f"dim_q={model.dim_q}), {n} chains per GPU, synthetic data; extra "
                                   "workload outside the BASELINE configs (SURVEY 8f rank 3)",
                       "unit_of_work": "one constrained leapfrog step: two half kicks with dh1_dpos "
                                       "(gradient of log_det_sqrt_gram) + cotangent projections, "
                                       "projection onto the manifold, reversibility check",
                       "step_size": dt, "solver": args.solver, "chains_total": n * world,
                       "ok_fraction": ok_frac,
                       "mean_projection_iters": [float(out["iters_fwd"].double().mean()),
                                                 float(out["iters_rev"].double().mean())],
                       "l2": "working set (blocks + scratch, > 1 GB) larger than L2"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach / hbm_peak, "traffic": None, "kernel": kname,
                         "peak_source": src, "ms_per_launch": kms,
                         "algorithmic_bytes_per_chain": 8.0 * rows * dim_y},
            "e2e": {"value": e2e_value, "unit": "steps/s",
                    "h2d_bytes_per_step": int(2 * qh.numel() * 8),
                    "d2h_bytes_per_step": int(2 * qh.numel() * 8 + 4 * n),
                    "api": "StochasticVolatilityModel.leapfrog_step (pinned host arrays)"},
            "gpu_launches": launches, "clocks": clk, "cpu_baseline": cpu,
        })  # fmt: skip
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print to fd 1 (e.g. NCCL's
    # version banner) is diverted to stderr for the duration of the run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lambda obj: lines.append(json.dumps(obj))  # noqa: E731
    try:
        (run_reference(a, emit) if a.impl == "reference"
         else run_sv(a, emit) if a.workload == "sv" else run_mlb(a, emit))
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
