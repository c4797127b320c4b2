"""Multi-rank host logic on CPU (gloo, world_size 2): chain sharding and the cross-chain
reductions of SURVEY 8e give the same numbers as one process over all chains."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def ar1(n_chain, n_draw, rho, seed):
    rng = np.random.default_rng(seed)
    x = np.empty((n_chain, n_draw))
    x[:, 0] = rng.standard_normal(n_chain)
    e = rng.standard_normal((n_chain, n_draw)) * np.sqrt(1 - rho * rho)
    for t in range(1, n_draw):
        x[:, t] = rho * x[:, t - 1] + e[:, t]
    return x + 0.05 * rng.standard_normal((n_chain, 1))


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))  # fmt: skip
    from manifold_lifting_b200 import parallel as par

    par.init_from_env("gloo")
    first, count = par.shard_chains(n_total)
    x = torch.as_tensor(ar1(n_total, 400, 0.6, 1))[first : first + count]
    res = {
        "shard": (first, count),
        "rhat_rank": par.rhat(x, "rank"), "rhat_split": par.rhat(x, "split"),
        "ess_bulk": par.ess(x, "bulk"), "ess_mean": par.ess(x, "mean"),
        "summary": par.summary({"x": x})["x"],
        "counts": par.global_counts(failed=rank + 1, steps=10),
        "gathered": par.all_gather_chains(x[:, :3]).numpy(),
    }  # fmt: skip
    # the step-size finalisation reduction (adapters.DualAveragingStepSizeAdapter.finalize)
    log_eps = torch.log(torch.linspace(0.1, 0.2, n_total, dtype=torch.float64))[first : first + count]
    tot = torch.stack([torch.exp(log_eps).sum(), torch.tensor(float(count), dtype=torch.float64)])
    par.all_reduce_sum_(tot)
    res["step_size"] = float(tot[0] / tot[1])
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_chains_partitions_exactly():
    from manifold_lifting_b200.parallel import shard_chains

    for n, w in [(65536, 8), (10, 3), (7, 8), (8192, 8), (1, 2)]:
        blocks = [shard_chains(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f1 == f0 + c0
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_diagnostics_single_process_sanity():
    """AR(1): ESS ~ N (1-rho)/(1+rho); R-hat ~ 1 for stationary chains and > 1.1 when one
    chain is shifted; rank-normalised R-hat against a NumPy/SciPy restatement of
    Vehtari et al. (2021)."""
    from scipy import stats

    from manifold_lifting_b200 import parallel as par

    x = ar1(8, 2000, 0.5, 0)
    e = par.ess(torch.as_tensor(x), "mean")
    assert 0.75 < e / (8 * 2000 * (1 - 0.5) / (1 + 0.5)) < 1.25
    assert abs(par.rhat(torch.as_tensor(x), "rank") - 1.0) < 0.01
    bad = x.copy()
    bad[0] += 3.0
    assert par.rhat(torch.as_tensor(bad), "rank") > 1.1

    def np_rhat(a):
        n = a.shape[1] // 2
        a = np.concatenate([a[:, :n], a[:, a.shape[1] - n:]], 0)

        def zs(b):
            r = stats.rankdata(b, method="average").reshape(b.shape)
            return stats.norm.ppf((r - 0.375) / (b.size + 0.25))

        def rh(b):
            w = b.var(1, ddof=1).mean()
            vp = (n - 1) / n * w + b.mean(1).var(ddof=1)
            return np.sqrt(vp / w)

        return max(rh(zs(a)), rh(zs(np.abs(a - np.median(a)))))

    assert abs(par.rhat(torch.as_tensor(bad), "rank") - np_rhat(bad)) < 1e-10
    assert abs(par.rhat(torch.as_tensor(x), "rank") - np_rhat(x)) < 1e-10


@pytest.mark.timeout(300)
def test_two_rank_reductions_match_single_process():
    from manifold_lifting_b200 import parallel as par

    n_total, world = 11, 2  # uneven split on purpose
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q))
             for r in range(world)]  # fmt: skip
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    x = torch.as_tensor(ar1(n_total, 400, 0.6, 1))
    want = {"rhat_rank": par.rhat(x, "rank"), "rhat_split": par.rhat(x, "split"),
            "ess_bulk": par.ess(x, "bulk"), "ess_mean": par.ess(x, "mean")}  # fmt: skip
    assert [got[r]["shard"] for r in range(world)] == [(0, 6), (6, 5)]
    for r in range(world):
        for k, v in want.items():
            assert abs(got[r][k] - v) < 1e-9 * max(1.0, abs(v)), (k, got[r][k], v)
        s, s1 = got[r]["summary"], par.summary({"x": x})["x"]
        for k in s1:
            assert abs(s[k] - s1[k]) < 1e-9 * max(1.0, abs(s1[k])), k
        assert got[r]["counts"] == {"failed": 3.0, "steps": 20.0}
        assert np.array_equal(got[r]["gathered"], x[:, :3].numpy())
        assert abs(got[r]["step_size"] - float(torch.linspace(0.1, 0.2, n_total, dtype=torch.float64).mean())) < 1e-14
