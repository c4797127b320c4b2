"""Multi-GPU plumbing: chains sharded over ranks, cross-chain reductions over NCCL.

The reference parallelises over chains with one OS process per chain and no
communication (toy_2d_model.py:450, utils.py:449-453).  Here one process drives one GPU
and owns a contiguous block of chains; the leapfrog / projection path has NO collective.
Collectives (`torch.distributed`, NCCL on GPUs, gloo in the CPU tests) are used only for

  * the step-size finalisation of warm-up (mean over ALL chains of the per-chain
    dual-averaging step sizes; `adapters.DualAveragingStepSizeAdapter.finalize`),
  * convergence diagnostics: split-R-hat and effective sample size from per-chain
    sufficient statistics (sums of chain means / variances / autocovariances), i.e. a
    few KB per variable; the rank-normalised variants (the ArviZ 0.16 defaults the
    reference reports through `arviz.summary`, utils.py:509-518) additionally need the
    global ranks of the draws and all-gather the (thinned) traces,
  * global failure counters.

Everything here works on tensors of either device; results do not depend on the number
of ranks (tests/test_cpu_dist.py checks world_size 2 against a single process).
"""
import math
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun) and bind
    this process to cuda:LOCAL_RANK.  No-op for a single process."""
    n = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if n > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return world()


def bind_to_gpu_cpus(local_rank=None):
    """Pin this process (and the pinned host buffers it allocates afterwards, by first
    touch) to the CPU cores NVML reports as local to its GPU, so that with one process per
    GPU on a multi-socket host the host <-> device copies do not cross the socket
    interconnect.  (On this pool's 8 x B200 boxes NVML reports all 32 cores as local to
    every GPU, so it changes nothing there: the host-entry throughput of 8 ranks stays at
    4.2x of one rank - the ranks share PCIe uplinks - while the device throughput is 7.9x.)
    Returns the number of cores bound to, or 0 if NVML / affinity calls are unavailable."""
    if not hasattr(os, "sched_setaffinity"):
        return 0
    local = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else int(local_rank)
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = (int(visible.split(",")[local])
                if visible and visible.replace(",", "").isdigit() else local)  # fmt: skip
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001  (NVML missing, containers without affinity rights, ...)
        return 0


def shard_chains(n_chain_total, rank=None, world_size=None):
    """Contiguous block of global chain indices owned by `rank`: (first, count).  The
    first `n % world` ranks get one extra chain.  Global chain ids key the Philox streams,
    so the sampled chains do not depend on the sharding."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(int(n_chain_total), world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def all_reduce_sum_(t):
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_max_(t):
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def all_gather_chains(x):
    """Concatenate per-rank [n_local, ...] blocks along dim 0 in rank order (block sizes
    may differ by one, see `shard_chains`)."""
    r, w = world()
    if w == 1:
        return x
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    pad = max(sizes)
    buf = x.new_zeros((pad,) + tuple(x.shape[1:]))
    buf[: x.shape[0]] = x
    out = [torch.empty_like(buf) for _ in range(w)]
    dist.all_gather(out, buf.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def global_counts(**local):
    """Sum integer / float counters over ranks: global_counts(failed=3, steps=100)."""
    keys = sorted(local)
    t = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64,
                     device="cuda" if torch.cuda.is_available() and dist.is_initialized()
                     and dist.get_backend() == "nccl" else "cpu")  # fmt: skip
    all_reduce_sum_(t)
    return {k: float(v) for k, v in zip(keys, t)}


# ------------------------------------------------------------------- diagnostics


def _split(x):
    """[n_chain, n_draw] -> [2 n_chain, n_draw // 2] (ArviZ `_split_chains`)."""
    n = x.shape[1] // 2
    return torch.cat([x[:, :n], x[:, x.shape[1] - n :]], 0)


def _autocov_sum(x):
    """Sum over chains of the biased autocovariance (normalised by n_draw), via FFT."""
    n = x.shape[1]
    m = 1 << (2 * n - 1).bit_length()
    xc = x - x.mean(1, keepdim=True)
    f = torch.fft.rfft(xc, n=m, dim=1)
    ac = torch.fft.irfft(f * f.conj(), n=m, dim=1)[:, :n] / n
    return ac.sum(0)


def _sufficient(x):
    """Per-rank sums over local chains: [count, sum mean, sum mean^2, sum var(ddof=1)]
    and the summed autocovariance sequence."""
    mean = x.mean(1)
    stats = torch.stack([torch.tensor(float(x.shape[0]), dtype=x.dtype, device=x.device),
                         mean.sum(), (mean * mean).sum(), x.var(1, unbiased=True).sum()])  # fmt: skip
    return stats, _autocov_sum(x)


def _rhat_from(stats, n_draw):
    m, s1, s2, sv = (float(v) for v in stats)
    w = sv / m
    b_over_n = (s2 - s1 * s1 / m) / (m - 1.0)  # variance of chain means, ddof=1
    var_plus = (n_draw - 1.0) / n_draw * w + b_over_n
    return math.sqrt(var_plus / w)


def _ess_from(stats, acov_sum, n_draw):
    """Geyer initial-monotone-sequence ESS over all chains (ArviZ 0.16 `_ess`)."""
    m, s1, s2, _ = (float(v) for v in stats)
    acov = (acov_sum / m).tolist()
    mean_var = acov[0] * n_draw / (n_draw - 1.0)
    var_plus = mean_var * (n_draw - 1.0) / n_draw
    if m > 1:
        var_plus += (s2 - s1 * s1 / m) / (m - 1.0)
    rho = [0.0] * n_draw
    even, rho[0] = 1.0, 1.0
    odd = 1.0 - (mean_var - acov[1]) / var_plus
    rho[1] = odd
    t = 1
    while t < (n_draw - 3) and (even + odd) > 0.0:
        even = 1.0 - (mean_var - acov[t + 1]) / var_plus
        odd = 1.0 - (mean_var - acov[t + 2]) / var_plus
        if (even + odd) >= 0:
            rho[t + 1], rho[t + 2] = even, odd
        t += 2
    max_t = t - 2
    if even > 0:
        rho[max_t + 1] = even
    t = 1
    while t <= max_t - 2:
        if (rho[t + 1] + rho[t + 2]) > (rho[t - 1] + rho[t]):
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    ess = m * n_draw
    tau = -1.0 + 2.0 * sum(rho[: max_t + 1]) + sum(rho[max_t + 1 : max_t + 2])
    tau = max(tau, 1.0 / math.log10(ess))
    return ess / tau


def _reduced(x):
    stats, ac = _sufficient(x)
    all_reduce_sum_(stats), all_reduce_sum_(ac)
    return stats, ac


def _rank_normalise(x_local):
    """z-scores of the global average ranks (Vehtari et al. 2021, eq. 14), computed on the
    all-gathered draws; returns this rank's block."""
    full = all_gather_chains(x_local)
    flat = full.reshape(-1)
    order = torch.argsort(flat, stable=True)
    ranks = torch.empty_like(flat)
    ranks[order] = torch.arange(1, flat.numel() + 1, dtype=flat.dtype, device=flat.device)
    # average ranks over ties
    vals, inv, cnt = torch.unique(flat, return_inverse=True, return_counts=True)
    if vals.numel() != flat.numel():
        sums = torch.zeros_like(vals).scatter_add_(0, inv, ranks)
        ranks = (sums / cnt.to(flat.dtype))[inv]
    z = (ranks - 0.375) / (flat.numel() + 0.25)
    z = math.sqrt(2.0) * torch.erfinv(2.0 * z - 1.0)
    z = z.reshape(full.shape)
    first, count = 0, x_local.shape[0]
    r, w = world()
    if w > 1:
        n = torch.tensor([x_local.shape[0]], dtype=torch.int64, device=x_local.device)
        sizes = [torch.zeros_like(n) for _ in range(w)]
        dist.all_gather(sizes, n)
        first = sum(int(s) for s in sizes[:r])
    return z[first : first + count]


def rhat(x, method="rank"):
    """Split-R-hat of one scalar variable traced as [n_chain_local, n_draw] on every
    rank.  method "rank" = max of the rank-normalised bulk and folded-tail R-hat (ArviZ
    default), "split" = classical split-R-hat (pure all-reduce)."""
    x = _split(x.to(torch.float64))
    n_draw = x.shape[1]
    if method == "split":
        return _rhat_from(_reduced(x)[0], n_draw)
    med = all_gather_chains(x).median()
    bulk = _rhat_from(_reduced(_rank_normalise(x))[0], n_draw)
    tail = _rhat_from(_reduced(_rank_normalise((x - med).abs()))[0], n_draw)
    return max(bulk, tail)


def ess(x, method="bulk"):
    """Effective sample size over all ranks' chains.  "bulk" = rank-normalised split
    chains (ArviZ default), "mean" = split chains without rank normalisation (pure
    all-reduce of sufficient statistics)."""
    x = _split(x.to(torch.float64))
    n_draw = x.shape[1]
    if method == "bulk":
        x = _rank_normalise(x)
    stats, ac = _reduced(x)
    return _ess_from(stats, ac, n_draw)


def summary(traces, var_names=None):
    """{name: {mean, sd, ess_bulk, r_hat}} over all ranks, for scalar traces
    [n_chain_local, n_draw] (the columns of `arviz.summary` the reference stores,
    utils.py:509-518)."""
    out = {}
    for name in var_names or sorted(traces):
        x = traces[name].to(torch.float64)
        s = torch.stack([torch.tensor(float(x.numel()), dtype=x.dtype, device=x.device),
                         x.sum(), (x * x).sum()])  # fmt: skip
        all_reduce_sum_(s)
        n, s1, s2 = (float(v) for v in s)
        mean = s1 / n
        out[name] = {"mean": mean, "sd": math.sqrt(max(s2 / n - mean * mean, 0.0) * n / (n - 1)),
                     "ess_bulk": ess(x, "bulk"), "r_hat": rhat(x, "rank")}  # fmt: skip
    return out
