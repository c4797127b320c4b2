"""Run outputs (reference example_models/utils.py:438-463, 509-518): file layout and the
summary columns against their defining properties on draws with known answers.  ArviZ is
not installable here, so the columns are checked by what they must equal, not against it."""
import json
import os
import pickle

import numpy as np
import torch

from manifold_lifting_b200 import outputs


def test_summary_columns_on_independent_normal_draws(tmp_path):
    rng = np.random.default_rng(0)
    n_chain, n_draw = 8, 2000
    x = 3.0 + 2.0 * rng.standard_normal((n_chain, n_draw))
    row = outputs.summarize_variable(x)
    n = n_chain * n_draw
    assert abs(row["mean"] - 3.0) < 0.06 and abs(row["sd"] - 2.0) < 0.05
    # independent draws: every ESS is the sample size up to estimation noise, R-hat is 1
    for k in ("ess_bulk", "ess_tail"):
        assert 0.8 * n < row[k] < 1.25 * n, (k, row[k])
    assert abs(row["r_hat"] - 1.0) < 0.01
    assert abs(row["mcse_mean"] - 2.0 / np.sqrt(n)) < 0.2 * 2.0 / np.sqrt(n)
    # sd of the sample sd of normal draws is sd / sqrt(2 n)
    assert abs(row["mcse_sd"] - 2.0 / np.sqrt(2 * n)) < 0.25 * 2.0 / np.sqrt(2 * n)
    # 94 % interval of N(3, 2): +-1.8808 sd, and it is the narrowest such interval
    assert abs(row["hdi_3%"] - (3.0 - 1.8808 * 2.0)) < 0.15
    assert abs(row["hdi_97%"] - (3.0 + 1.8808 * 2.0)) < 0.15
    inside = ((x >= row["hdi_3%"]) & (x <= row["hdi_97%"])).mean()
    assert abs(inside - 0.94) < 2e-3


def test_summary_detects_autocorrelation_and_disagreeing_chains():
    rng = np.random.default_rng(1)
    n_chain, n_draw, rho = 4, 4000, 0.9
    e = rng.standard_normal((n_chain, n_draw))
    x = np.empty_like(e)
    x[:, 0] = e[:, 0]
    for t in range(1, n_draw):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho**2) * e[:, t]
    row = outputs.summarize_variable(x)
    want = n_chain * n_draw * (1 - rho) / (1 + rho)  # AR(1): tau = (1 + rho) / (1 - rho)
    assert 0.6 * want < row["ess_bulk"] < 1.5 * want
    assert row["mcse_mean"] > 3.0 * row["sd"] / np.sqrt(n_chain * n_draw)
    shifted = x + np.arange(n_chain)[:, None]  # chains centred at 0, 1, 2, 3
    assert outputs.summarize_variable(shifted)["r_hat"] > 1.5


def test_files_written_like_the_reference_harness(tmp_path):
    import manifold_lifting_b200 as mlb

    rng = np.random.default_rng(2)
    traces = {"u": rng.standard_normal((3, 50, 2)), "tau": rng.standard_normal((3, 50)),
              "constr_calls": np.cumsum(np.ones((3, 50), dtype=np.int64), 1)}  # fmt: skip
    stats = {"accept_stat": rng.uniform(size=3)}
    out = str(tmp_path)
    paths = outputs.save_traces(out, traces, stats)
    assert len(paths) == 3 * 3 + 3
    # Mici nests monitored statistics per transition: stats_{chain}_{transition}_{name}.npy
    nested = outputs.save_traces(out, {}, {"integration": {"accept_stat": rng.uniform(size=(3, 50)),
                                                          "convergence_error": np.zeros((3, 50))}})
    assert sorted(os.path.basename(p) for p in nested)[:2] == [
        "stats_0_integration_accept_stat.npy", "stats_0_integration_convergence_error.npy"]
    back = outputs.load_traces(out, ["u", "tau"], 3)
    assert isinstance(back["u"][1], np.memmap) and np.array_equal(back["u"][1], traces["u"][1])
    rows, summ = outputs.compute_and_save_summary(out, ["u", "tau"], traces,
                                                  total_sampling_time=1.5)  # fmt: skip
    on_disk = json.load(open(tmp_path / "summary.json"))
    assert set(on_disk["mean"]) == {"u[0]", "u[1]", "tau"}
    assert on_disk["total_sampling_time"] == 1.5
    assert on_disk["per_chain_constr_calls"] == [50, 50, 50]
    assert set(on_disk) >= {"mean", "sd", "hdi_3%", "hdi_97%", "mcse_mean", "mcse_sd",
                            "ess_bulk", "ess_tail", "r_hat"}  # fmt: skip

    class State:  # stands in for states.ChainState (needs a CUDA device to construct)
        pos, mom, dir = torch.zeros(3, 4), torch.ones(3, 4), 1

    outputs.save_final_states(out, State)
    states = pickle.load(open(tmp_path / "final_chain_states.pkl", "rb"))
    assert len(states) == 3 and states[2]["mom"].shape == (4,) and states[0]["dir"] == 1
    assert hasattr(mlb, "outputs")
