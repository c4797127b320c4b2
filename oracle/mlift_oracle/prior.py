"""Oracle restatement of the prior machinery used by the four BASELINE models.

Follows (all paths under /root/reference/):
  * mlift/prior.py:17-118        prior specification -> generate_params / neg log dens
  * mlift/distributions.py:66-106 pull-back of a density through a bounding transform
  * mlift/distributions.py:141-203, 439-465  normal / log_normal / half_cauchy
  * mlift/transforms.py:16-55, 87-95, 137-144  exp-bounding, affine, N(0,1)->half-Cauchy
  * mlift/math.py:22-29          standard Cauchy (i)cdf

Only the slices the configs need are restated (normal, log-normal, half-Cauchy priors;
"unbounded" and "normal" parametrisations).  Everything is torch fp64 so that
`torch.func` transforms can stand in for the JAX transforms the reference applies.
TEST INFRASTRUCTURE ONLY - see package docstring.
"""
import math
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np
import torch

_INF = float("inf")


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)


def ndtr(x):
    return torch.special.ndtr(_t(x))


@dataclass
class Transform:
    """Elementwise monotone map with derivative (transforms.py:16-40)."""

    forward: Callable
    dforward: Callable  # derivative of forward, elementwise
    domain: Tuple[float, float]
    image: Tuple[float, float]

    def forward_and_det_jacobian(self, u):
        # transforms.py:32-37: scalar -> (x, dx/du); vector -> (x, SUM of dx/du)
        x, dx = self.forward(u), self.dforward(u)
        if dx.ndim == 0:
            return x, dx
        return x, dx.sum()


def unbounded_to_lower_bounded(lower):
    # transforms.py:43-55
    return Transform(
        forward=lambda u: torch.exp(_t(u)) + lower,
        dforward=lambda u: torch.exp(_t(u)),
        domain=(-_INF, _INF),
        image=(lower, _INF),
    )


def diagonal_affine_map(location, scale):
    # transforms.py:87-95
    return Transform(
        forward=lambda x: location + scale * _t(x),
        dforward=lambda x: scale * torch.ones_like(_t(x)),
        domain=(-_INF, _INF),
        image=(-_INF, _INF),
    )


def standard_cauchy_icdf(u):
    # math.py:27-29
    return torch.tan(math.pi * (u - 0.5))


def standard_normal_to_half_cauchy(scale):
    # transforms.py:137-144
    def fwd(n):
        return standard_cauchy_icdf((ndtr(n) + 1) / 2) * scale

    def dfwd(n):
        n = _t(n)
        phi = torch.exp(-0.5 * n * n) / math.sqrt(2 * math.pi)
        arg = math.pi * ((ndtr(n) + 1) / 2 - 0.5)
        return scale * (math.pi / 2) * phi / torch.cos(arg) ** 2

    return Transform(fwd, dfwd, (-_INF, _INF), (0.0, _INF))


def standard_normal_to_log_normal(location, scale):
    # distributions.py:190-195
    return Transform(
        forward=lambda n: torch.exp(location + scale * _t(n)),
        dforward=lambda n: scale * torch.exp(location + scale * _t(n)),
        domain=(-_INF, _INF),
        image=(0.0, _INF),
    )


@dataclass
class Distribution:
    """distributions.py:22-62 (density part only; sampling is host NumPy)."""

    nld_elem: Callable  # elementwise negative log density (unnormalised)
    sample: Callable  # (rng, shape) -> ndarray
    support: Tuple[float, float]
    from_standard_normal_transform: Optional[Transform] = None

    def neg_log_dens(self, x):
        nld = self.nld_elem(x)
        return nld.sum() if nld.ndim > 0 else nld


def normal(location, scale):
    # distributions.py:141-168
    return Distribution(
        nld_elem=lambda x: ((_t(x) - location) / scale) ** 2 / 2,
        sample=lambda rng, shape=(): rng.normal(loc=location, scale=scale, size=shape),
        support=(-_INF, _INF),
        from_standard_normal_transform=diagonal_affine_map(location, scale),
    )


def log_normal(location, scale):
    # distributions.py:171-203
    return Distribution(
        nld_elem=lambda x: ((torch.log(_t(x)) - location) / scale) ** 2 / 2
        + torch.log(_t(x)),
        sample=lambda rng, shape=(): np.exp(
            rng.normal(loc=location, scale=scale, size=shape)
        ),
        support=(0.0, _INF),
        from_standard_normal_transform=standard_normal_to_log_normal(location, scale),
    )


def half_cauchy(scale):
    # distributions.py:439-465
    return Distribution(
        nld_elem=lambda x: torch.log1p((_t(x) / scale) ** 2),
        sample=lambda rng, shape=(): abs(rng.standard_cauchy(size=shape) * scale),
        support=(0.0, _INF),
        from_standard_normal_transform=standard_normal_to_half_cauchy(scale),
    )


def pullback_distribution(dist, transform):
    # distributions.py:66-106
    def nld(u):
        x, det = transform.forward_and_det_jacobian(u)
        return dist.neg_log_dens(x) - torch.log(det)

    def sample(rng, shape=()):
        x = dist.sample(rng, shape)
        assert transform.image[0] == 0.0  # only exp-type bounding is restated
        return np.log(x - transform.image[0])

    assert dist.support == transform.image
    out = Distribution(nld_elem=None, sample=sample, support=transform.domain)
    out.neg_log_dens = nld  # already summed
    return out


@dataclass
class PriorSpecification:
    # prior.py:10-14
    distribution: Distribution
    shape: Tuple[int, ...] = ()
    transform: Optional[Callable] = None


def _to_unbounded(spec):
    # prior.py:17-42 (only the lower-bounded and unbounded cases occur in the configs)
    lo, hi = spec.distribution.support
    if lo != -_INF and hi != _INF:
        raise NotImplementedError("two-sided bounded priors are outside the configs")
    if lo != -_INF:
        t = unbounded_to_lower_bounded(lo)
    elif hi != _INF:
        raise NotImplementedError("upper-bounded priors are outside the configs")
    else:
        return spec
    return PriorSpecification(
        distribution=pullback_distribution(spec.distribution, t),
        shape=spec.shape,
        transform=t.forward,
    )


def _to_standard_normal(spec):
    # prior.py:45-55
    t = spec.distribution.from_standard_normal_transform
    return PriorSpecification(
        distribution=normal(0.0, 1.0), shape=spec.shape, transform=t.forward
    )


def set_up_prior(prior_specs):
    """prior.py:58-118.  Returns (compute_dim_u, generate_params, prior_nld, sample)."""

    def respecs(data):
        for name, spec in prior_specs.items():
            if (
                data.get("parametrization") == "normal"
                and spec.distribution.from_standard_normal_transform is not None
            ):
                yield name, _to_standard_normal(spec)
            else:
                yield name, _to_unbounded(spec)

    def slices(u, data):
        i = 0
        for name, spec in respecs(data):
            size = int(np.prod(spec.shape))
            yield name, spec, (u[i] if spec.shape == () else u[i : i + size])
            i += size

    def compute_dim_u(data):
        return sum(int(np.prod(spec.shape)) for _, spec in respecs(data))

    def generate_params(u, data):
        return {
            name: (spec.transform(us) if spec.transform is not None else us)
            for name, spec, us in slices(u, data)
        }

    def prior_neg_log_dens(u, data):
        nld = 0.0
        for _, spec, us in slices(u, data):
            nld = nld + spec.distribution.neg_log_dens(us)
        return nld

    def sample_from_prior(rng, data):
        parts = [
            np.atleast_1d(np.asarray(spec.distribution.sample(rng, spec.shape))).ravel()
            for _, spec in respecs(data)
        ]
        return np.concatenate(parts)

    return compute_dim_u, generate_params, prior_neg_log_dens, sample_from_prior
