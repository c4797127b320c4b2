"""Oracle restatement of the four BASELINE model definitions (torch fp64).

Each model exposes `generate_y`, `extended_prior_neg_log_dens`, `compute_dim_u`,
`sample_from_prior` with the argument conventions of the reference so that the generic
system classes in `systems.py` can apply autodiff transforms to them exactly as the
reference applies JAX transforms.

Reference sources (under /root/reference/):
  * toy            mlift/example_models/toy_2d_model.py:43-87 and
                   notebooks/Two-dimensional_example.ipynb (factor-2 forward function)
  * eight schools  mlift/example_models/eight_schools.py:28-81
  * FitzHugh-Nagumo mlift/example_models/fitzhugh_nagumo.py:20-64
  * Hodgkin-Huxley  mlift/example_models/hh_current_stimulus.py:19-203
TEST INFRASTRUCTURE ONLY - see package docstring.
"""
import numpy as np
import torch

from .ode import integrate_ode_rk4
from .prior import PriorSpecification, half_cauchy, log_normal, normal, set_up_prior


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)


# --------------------------------------------------------------------------- toy 2-D


class Toy2D:
    """q = (theta0, theta1, eta); c(q) = F(theta) + sigma*eta - y.

    `coeff`=2 is the notebook's forward function (BASELINE configs 1-2), `coeff`=1 the
    script's (toy_2d_model.py:45).  Prior nld is 0.5*|q|^2 (toy_2d_model.py:72-74).
    """

    dim_theta, dim_y = 2, 1

    def __init__(self, sigma, y=1.0, coeff=2.0):
        self.sigma, self.y, self.coeff = float(sigma), float(y), float(coeff)

    def forward_func(self, theta):
        return theta[1] ** 2 + self.coeff * theta[0] ** 2 * (theta[0] ** 2 - 0.5)

    def constr(self, q):
        return (self.forward_func(q[:2]) + self.sigma * q[2] - self.y).reshape(1)

    @staticmethod
    def neg_log_dens(q):
        return (q**2).sum() / 2

    def lift(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        f = theta[..., 1] ** 2 + self.coeff * theta[..., 0] ** 2 * (
            theta[..., 0] ** 2 - 0.5
        )
        return np.concatenate([theta, ((self.y - f) / self.sigma)[..., None]], -1)


# --------------------------------------------------------------------- eight schools


class EightSchools:
    """y = mu + tau*v + sigma*n (eight_schools.py:28-53)."""

    prior_specifications = {
        "mu": PriorSpecification(distribution=normal(0.0, 5.0)),
        "tau": PriorSpecification(distribution=half_cauchy(5.0)),
    }

    def __init__(self, y_obs, sigma, parametrization="unbounded"):
        self.data = {
            "y_obs": _t(np.asarray(y_obs, dtype=np.float64)),
            "sigma": _t(np.asarray(sigma, dtype=np.float64)),
            "parametrization": parametrization,
        }
        (
            self._compute_dim_u,
            self.generate_params,
            self.prior_neg_log_dens,
            self._sample_from_prior,
        ) = set_up_prior(self.prior_specifications)
        self.dim_u = self._compute_dim_u(self.data)
        self.dim_y = int(self.data["y_obs"].shape[0])

    def generate_from_model(self, u, v, data):
        params = self.generate_params(u, data)
        return params, params["mu"] + params["tau"] * v

    def generate_y(self, u, v, n, data):
        _, x = self.generate_from_model(u, v, data)
        return x + data["sigma"] * n

    def extended_prior_neg_log_dens(self, q):
        U, Y = self.dim_u, self.dim_y
        u, v, n = q[:U], q[U : U + Y], q[U + Y :]
        return (
            self.prior_neg_log_dens(u, self.data) + (v**2).sum() / 2 + (n**2).sum() / 2
        )

    def sample_initial_states(self, rng, num_chain):
        """eight_schools.py:67-81: u~prior, v~N(0,I), n=(y-x)/sigma (on manifold)."""
        out = []
        for _ in range(num_chain):
            u = self._sample_from_prior(rng, self.data)
            v = rng.standard_normal(self.dim_y)
            with torch.no_grad():
                _, x = self.generate_from_model(_t(u), _t(v), self.data)
            n = (self.data["y_obs"] - x) / self.data["sigma"]
            out.append(np.concatenate((u, v, n.numpy())))
        return np.stack(out)


# ------------------------------------------------------------------- FitzHugh-Nagumo


class FitzHughNagumo:
    """RK4-integrated 2-state ODE observed through x0 (fitzhugh_nagumo.py:20-64)."""

    prior_specifications = {
        "alpha": PriorSpecification(distribution=log_normal(0.0, 1.0)),
        "beta": PriorSpecification(distribution=log_normal(0.0, 1.0)),
        "gamma": PriorSpecification(distribution=log_normal(0.0, 1.0)),
        "delta": PriorSpecification(distribution=log_normal(-1.0, 1.0)),
        "epsilon": PriorSpecification(distribution=log_normal(-3.0, 1.0)),
        "zeta": PriorSpecification(distribution=log_normal(-2.0, 1.0)),
        "sigma": PriorSpecification(distribution=log_normal(-1.0, 1.0)),
        "x_init": PriorSpecification(shape=(2,), distribution=normal(0.0, 1.0)),
    }

    def __init__(self, y_obs, t_seq, dt, parametrization="unbounded"):
        self.data = {
            "y_obs": _t(np.asarray(y_obs, dtype=np.float64)),
            "t_seq": np.asarray(t_seq, dtype=np.float64),
            "dt": float(np.asarray(dt).reshape(-1)[0]),
            "parametrization": parametrization,
        }
        (
            self._compute_dim_u,
            self.generate_params,
            self.prior_neg_log_dens,
            self._sample_from_prior,
        ) = set_up_prior(self.prior_specifications)
        self.dim_u = self._compute_dim_u(self.data)
        self.dim_y = int(self.data["y_obs"].shape[0])

    @staticmethod
    def dx_dt(x, t, p):
        # fitzhugh_nagumo.py:35-42
        return torch.stack(
            (
                p["alpha"] * x[0] - p["beta"] * x[0] ** 3 + p["gamma"] * x[1],
                -p["delta"] * x[0] - p["epsilon"] * x[1] + p["zeta"],
            )
        )

    def generate_from_model(self, u, data):
        params = self.generate_params(u, data)
        x_seq = integrate_ode_rk4(
            self.dx_dt, params["x_init"], data["t_seq"], params, data["dt"]
        )
        return params, x_seq

    def generate_y(self, u, n, data):
        params, x = self.generate_from_model(u, data)
        return x[1:, 0] + params["sigma"] * n

    def extended_prior_neg_log_dens(self, q):
        u, n = q[: self.dim_u], q[self.dim_u :]
        return self.prior_neg_log_dens(u, self.data) + (n**2).sum() / 2

    def lift(self, u):
        """On-manifold q for a given u: n = (y_obs - x(u)) / sigma(u)."""
        with torch.no_grad():
            params, x = self.generate_from_model(_t(u), self.data)
            n = (self.data["y_obs"] - x[1:, 0]) / params["sigma"]
        return np.concatenate((np.asarray(u, dtype=np.float64), n.numpy()))


# ------------------------------------------------------------------- Hodgkin-Huxley

HH_CONSTANT_KEYS = (
    "k_alpha_n_1", "k_alpha_n_2", "k_alpha_n_3", "k_beta_n_1", "k_beta_n_2",
    "k_beta_n_3", "k_alpha_m_1", "k_alpha_m_2", "k_alpha_m_3", "k_beta_m_1",
    "k_beta_m_2", "k_beta_m_3", "k_alpha_h_1", "k_alpha_h_2", "k_alpha_h_3",
    "k_beta_h_1", "k_beta_h_2", "k_beta_h_3", "k_p_infinity_1", "k_p_infinity_2",
    "k_tau_p_2", "k_tau_p_3", "k_tau_p_4", "E_Na", "E_K", "E_leak", "c_m",
    "stimulus_start_time", "stimulus_end_time", "end_time", "i_stimulus", "dt",
    "observation_interval",
)  # fmt: skip


class HodgkinHuxley:
    """5-state conductance model, exponential-Euler steps (hh_current_stimulus.py)."""

    prior_specifications = {
        "k_tau_p_1": PriorSpecification(distribution=log_normal(8.0, 1.0)),
        "g_bar_Na": PriorSpecification(distribution=log_normal(4.0, 1.0)),
        "g_bar_K": PriorSpecification(distribution=log_normal(2.0, 1.0)),
        "g_bar_M": PriorSpecification(distribution=log_normal(-3.0, 1.0)),
        "g_leak": PriorSpecification(distribution=log_normal(-3.0, 1.0)),
        "v_t": PriorSpecification(distribution=normal(-60.0, 10.0)),
        "sigma": PriorSpecification(distribution=log_normal(0.0, 1.0)),
    }

    def __init__(self, y_obs, constants, parametrization="unbounded"):
        self.data = {k: float(np.asarray(constants[k])) for k in HH_CONSTANT_KEYS}
        self.data["y_obs"] = _t(np.asarray(y_obs, dtype=np.float64))
        self.data["parametrization"] = parametrization
        (
            self._compute_dim_u,
            self.generate_params,
            self.prior_neg_log_dens,
            self._sample_from_prior,
        ) = set_up_prior(self.prior_specifications)
        self.dim_u = self._compute_dim_u(self.data)
        self.dim_y = int(self.data["y_obs"].shape[0])

    # rate functions: hh_current_stimulus.py:34-114
    @staticmethod
    def _x_over_expm1_x(x):
        return x / torch.expm1(x)

    def _alpha_n(self, v, p):
        return (
            p["k_alpha_n_1"]
            * self._x_over_expm1_x(-(v - p["v_t"] - p["k_alpha_n_2"]) / p["k_alpha_n_3"])
            * p["k_alpha_n_3"]
        )

    @staticmethod
    def _beta_n(v, p):
        return p["k_beta_n_1"] * torch.exp(
            -(v - p["v_t"] - p["k_beta_n_2"]) / p["k_beta_n_3"]
        )

    def _alpha_m(self, v, p):
        return (
            p["k_alpha_m_1"]
            * self._x_over_expm1_x(-(v - p["v_t"] - p["k_alpha_m_2"]) / p["k_alpha_m_3"])
            * p["k_alpha_m_3"]
        )

    def _beta_m(self, v, p):
        return (
            p["k_beta_m_1"]
            * self._x_over_expm1_x((v - p["v_t"] - p["k_beta_m_2"]) / p["k_beta_m_3"])
            * p["k_beta_m_3"]
        )

    @staticmethod
    def _alpha_h(v, p):
        return p["k_alpha_h_1"] * torch.exp(
            -(v - p["v_t"] - p["k_alpha_h_2"]) / p["k_alpha_h_3"]
        )

    @staticmethod
    def _beta_h(v, p):
        return p["k_beta_h_1"] / (
            torch.exp(-(v - p["v_t"] - p["k_beta_h_2"]) / p["k_beta_h_3"]) + 1
        )

    @staticmethod
    def _p_infinity(v, p):
        return 1 / (1 + torch.exp((-v + p["k_p_infinity_1"]) / p["k_p_infinity_2"]))

    @staticmethod
    def _tau_p(v, p):
        return p["k_tau_p_1"] / (
            p["k_tau_p_2"] * torch.exp((v + p["k_tau_p_3"]) / p["k_tau_p_4"])
            + torch.exp(-(v + p["k_tau_p_3"]) / p["k_tau_p_4"])
        )

    def _x_init(self, p):
        # hh_current_stimulus.py:125-139: steady state at E_leak
        v = _t(p["E_leak"])
        an, bn = self._alpha_n(v, p), self._beta_n(v, p)
        am, bm = self._alpha_m(v, p), self._beta_m(v, p)
        ah, bh = self._alpha_h(v, p), self._beta_h(v, p)
        return torch.stack(
            (v + 0 * an, an / (an + bn), am / (am + bm), ah / (ah + bh),
             self._p_infinity(v, p) + 0 * an)
        )  # fmt: skip

    def time_grid(self):
        d = self.data
        n_step = int(d["end_time"] / d["dt"]) + 1
        t_seq = np.linspace(0, d["end_time"], n_step)
        return t_seq[:-1], t_seq[1:] - t_seq[:-1]

    def obs_slice(self):
        d = self.data
        return slice(
            int(d["stimulus_start_time"] / d["dt"]),
            int(d["stimulus_end_time"] / d["dt"]) + 1,
            int(d["observation_interval"] / d["dt"]),
        )

    def _integrate(self, x_init, p):
        # hh_current_stimulus.py:142-178
        t0s, dts = self.time_grid()
        x = x_init
        xs = [x_init]
        for t, dt in zip(t0s, dts):
            v, n, m, h, pp = x
            g_K = p["g_bar_K"] * n**4
            g_Na = p["g_bar_Na"] * m**3 * h
            g_M = p["g_bar_M"] * pp
            tau_v = p["c_m"] / (g_K + g_Na + g_M + p["g_leak"])
            stim = (
                0.0
                if (t < p["stimulus_start_time"] or t > p["stimulus_end_time"])
                else p["i_stimulus"]
            )
            v_inf = (
                (g_K * p["E_K"] + g_Na * p["E_Na"] + g_M * p["E_K"]
                 + p["g_leak"] * p["E_leak"] + stim)
                * tau_v / p["c_m"]
            )  # fmt: skip
            v_ = v_inf + (v - v_inf) * torch.exp(-dt / tau_v)
            an, bn = self._alpha_n(v_, p), self._beta_n(v_, p)
            am, bm = self._alpha_m(v_, p), self._beta_m(v_, p)
            ah, bh = self._alpha_h(v_, p), self._beta_h(v_, p)
            tau_n, n_inf = 1 / (an + bn), an / (an + bn)
            tau_m, m_inf = 1 / (am + bm), am / (am + bm)
            tau_h, h_inf = 1 / (ah + bh), ah / (ah + bh)
            n_ = n_inf + (n - n_inf) * torch.exp(-dt / tau_n)
            m_ = m_inf + (m - m_inf) * torch.exp(-dt / tau_m)
            h_ = h_inf + (h - h_inf) * torch.exp(-dt / tau_h)
            p_inf = self._p_infinity(v_, p)
            p_ = p_inf + (pp - p_inf) * torch.exp(-dt / self._tau_p(v_, p))
            x = torch.stack((v_, n_, m_, h_, p_))
            xs.append(x)
        return torch.stack(xs)

    def generate_from_model(self, u, data):
        params = self.generate_params(u, data)
        pd = {**data, **params}
        x_seq = self._integrate(self._x_init(pd), pd)
        return params, x_seq[self.obs_slice(), 0]

    def generate_y(self, u, n, data):
        params, x = self.generate_from_model(u, data)
        return x + params["sigma"] * n

    def extended_prior_neg_log_dens(self, q):
        u, n = q[: self.dim_u], q[self.dim_u :]
        return self.prior_neg_log_dens(u, self.data) + (n**2).sum() / 2

    def lift(self, u):
        with torch.no_grad():
            params, x = self.generate_from_model(_t(u), self.data)
            n = (self.data["y_obs"] - x) / params["sigma"]
        return np.concatenate((np.asarray(u, dtype=np.float64), n.numpy()))
