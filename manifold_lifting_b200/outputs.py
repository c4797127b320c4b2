"""Files the reference's experiment harness leaves behind for one run (reference
mlift/example_models/utils.py:438-463 `sample_chains`, :509-518
`compute_and_save_summary`): per-chain trace / statistic arrays as `.npy` files that can
be memory-mapped, `final_chain_states.pkl`, and `summary.json` = the `arviz.summary`
table as a dict of columns plus run-level entries.

What is restated and what is not verifiable here: ArviZ and Mici are absent from this image
(SURVEY.md 8c).  The summary columns follow ArviZ 0.16's definitions as published (mean,
sd, hdi_3% / hdi_97% = narrowest 94 % interval of the pooled draws, mcse_mean, mcse_sd,
ess_bulk, ess_tail, r_hat; `parallel.py` holds the ESS / R-hat estimators) and are checked
in tests/test_cpu_outputs.py against their defining properties on draws with known
answers.  The per-chain file names are Mici 0.2.1's (`force_memmap=True,
memmap_path=output_dir`, utils.py:449-450): `trace_{chain}_{name}.npy` and, for the
statistics Mici nests per transition (`stats["integration"]["accept_stat"]`,
monitor_stats of utils.py:335-340), `stats_{chain}_{transition}_{name}.npy` - as recalled
from Mici's `_generate_memmap_filename` (its source is not in the reference tree).  Nothing
in the reference READS these files (grep: the harness's own products are `summary.json`,
`final_chain_states.pkl` and, for the Euclidean baseline only, `final_metric.npy`), so the
names matter to users' own post-processing only.
"""
import json
import math
import os
import pickle

import numpy as np
import torch

from . import parallel


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save_traces(output_dir, traces, stats=None, chain_offset=0):
    """traces[name]: [n_chain, n_draw, ...]; stats[name]: [n_chain] or [n_chain, n_draw].
    One `.npy` per chain and name (`np.load(..., mmap_mode="r")` maps them)."""
    os.makedirs(output_dir, exist_ok=True)
    paths = []
    flat_stats = {}
    for name, arr in (stats or {}).items():
        if isinstance(arr, dict):  # Mici nests statistics per transition
            flat_stats.update({f"{name}_{k}": v for k, v in arr.items()})
        elif isinstance(name, tuple):
            flat_stats["_".join(name)] = arr
        else:
            flat_stats[name] = arr
    for kind, group in (("trace", traces), ("stats", flat_stats)):
        for name, arr in group.items():
            a = _np(arr)
            for i in range(a.shape[0]):
                p = os.path.join(output_dir, f"{kind}_{chain_offset + i}_{name}.npy")
                np.save(p, np.ascontiguousarray(a[i]))
                paths.append(p)
    return paths


def load_traces(output_dir, names, n_chain, kind="trace", chain_offset=0, mmap_mode="r"):
    return {name: [np.load(os.path.join(output_dir, f"{kind}_{chain_offset + i}_{name}.npy"),
                           mmap_mode=mmap_mode) for i in range(n_chain)]
            for name in names}  # fmt: skip


def save_final_states(output_dir, state):
    """`final_chain_states.pkl` (utils.py:455-456): a list with one entry per chain holding
    what a Mici `ChainState` carries across runs (pos, mom, dir) as NumPy arrays."""
    pos = _np(state.pos)
    mom = None if state.mom is None else _np(state.mom)
    states = [{"pos": pos[i].copy(), "mom": None if mom is None else mom[i].copy(),
               "dir": int(state.dir)} for i in range(pos.shape[0])]  # fmt: skip
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "final_chain_states.pkl"), mode="w+b") as f:
        pickle.dump(states, f)
    return states


def _ess_mean(x):
    """ESS of split chains without rank normalisation (ArviZ `_ess_mean`)."""
    xs = parallel._split(torch.as_tensor(x, dtype=torch.float64))
    stats, ac = parallel._sufficient(xs)
    return parallel._ess_from(stats, ac, xs.shape[1])


def hdi(x, prob=0.94):
    """Narrowest interval containing `prob` of the pooled draws (ArviZ `hdi`, unimodal)."""
    s = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = s.size
    k = int(math.floor(prob * n))
    width = s[k:] - s[: n - k]
    j = int(np.argmin(width))
    return float(s[j]), float(s[j + k])


def summarize_variable(x):
    """One row of `arviz.summary` for draws x [n_chain, n_draw]."""
    x = np.asarray(_np(x), dtype=np.float64)
    t = torch.as_tensor(x)
    sd = float(x.std(ddof=1))
    ess_mean = _ess_mean(x)
    ess_sd = min(ess_mean, _ess_mean((x - x.mean()) ** 2))
    fac = math.sqrt(math.e * (1.0 - 1.0 / ess_sd) ** (ess_sd - 1.0) - 1.0)
    q05, q95 = np.quantile(x, [0.05, 0.95])
    ess_tail = min(_ess_mean((x <= q05).astype(np.float64)),
                   _ess_mean((x <= q95).astype(np.float64)))  # fmt: skip
    lo, hi = hdi(x)
    return {"mean": float(x.mean()), "sd": sd, "hdi_3%": lo, "hdi_97%": hi,
            "mcse_mean": sd / math.sqrt(ess_mean), "mcse_sd": sd * fac,
            "ess_bulk": float(parallel.ess(t, "bulk")), "ess_tail": float(ess_tail),
            "r_hat": float(parallel.rhat(t, "rank"))}  # fmt: skip


def compute_and_save_summary(output_dir, var_names, traces, **kwargs):
    """utils.py:509-518: `summary.json` with the summary table as {column: {variable:
    value}} (`DataFrame.to_dict()`), the run-level keyword entries, and
    `per_chain_<name>` for every `*_calls` trace (last value of each chain).  Vector
    variables get one row per component named `name[i]` as ArviZ does."""
    rows = {}
    for name in var_names:
        a = _np(traces[name])
        if a.ndim == 2:
            rows[name] = summarize_variable(a)
        else:
            flat = a.reshape(a.shape[0], a.shape[1], -1)
            for j in range(flat.shape[2]):
                rows[f"{name}[{j}]"] = summarize_variable(flat[:, :, j])
    columns = ["mean", "sd", "hdi_3%", "hdi_97%", "mcse_mean", "mcse_sd", "ess_bulk",
               "ess_tail", "r_hat"]  # fmt: skip
    summary_dict = {c: {v: rows[v][c] for v in rows} for c in columns}
    summary_dict.update(kwargs)
    for key, value in traces.items():
        if key[-6:] == "_calls":
            summary_dict["per_chain_" + key] = [int(_np(v)[-1]) for v in value]
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "summary.json"), mode="w") as f:
        json.dump(summary_dict, f, ensure_ascii=False, indent=2)
    return rows, summary_dict
