"""Model specifications: what replaces the reference's `generate_y` closures.

In the reference a model is an arbitrary Python function `generate_y(u[, v], n, data)`
traced by JAX (mlift/systems.py:520-527, 662-669).  A Python closure cannot cross a C
ABI, so here a model is a `ModelSpec`: a model id + constant data blob handed to
`mlb_model_create`, with the forward map, its Jacobian blocks and the second-derivative
contraction needed for grad log det hand-written in CUDA (csrc/mlb_models.cuh).
A `ModelSpec` is passed wherever the reference takes `generate_y`.
"""
import ctypes as C

import numpy as np

from . import _lib


class ModelSpec:
    """Handle of a compiled-in model (wraps `mlb_model*`)."""

    family = None  # "additive" | "hierarchical"

    def __init__(self, model_id, dim_y, parametrization, data):
        self._data = np.ascontiguousarray(data, dtype=np.float64).ravel()
        par = {"unbounded": _lib.PARAM_UNBOUNDED, "normal": _lib.PARAM_NORMAL}[parametrization]
        desc = _lib.ModelDesc(model_id, dim_y, par, self._data.size,
                              self._data.ctypes.data_as(C.POINTER(C.c_double)))  # fmt: skip
        h = C.c_void_p()
        _lib.check(_lib.lib().mlb_model_create(C.byref(desc), C.byref(h)))
        self.handle = h
        L = _lib.lib()
        self.dim_u, self.dim_y, self.dim_q = (L.mlb_model_dim_u(h), L.mlb_model_dim_y(h),
                                              L.mlb_model_dim_q(h))  # fmt: skip
        self.jac_rows, self.gram_rows = L.mlb_model_jac_rows(h), L.mlb_model_gram_rows(h)
        self.gram_dim = L.mlb_model_gram_dim(h)
        self.parametrization = parametrization
        self.model_id = model_id

    def __del__(self):
        try:
            _lib.lib().mlb_model_destroy(self.handle)
        except Exception:
            pass

    def __call__(self, *args, **kwargs):
        raise TypeError("a ModelSpec stands in for `generate_y`; it is evaluated on the "
                        "GPU through the system classes, not called from Python")  # fmt: skip


class Toy2DModel(ModelSpec):
    """Toy model F(theta) = theta1^2 + coeff*theta0^2*(theta0^2 - 1/2), y = F + sigma*eta.

    Reference: mlift/example_models/toy_2d_model.py:43-87 (coeff=1) and
    notebooks/Two-dimensional_example.ipynb (coeff=2, the BASELINE config).
    """

    family = "additive"

    def __init__(self, sigma, y=1.0, coeff=2.0):
        super().__init__(_lib.MODEL_TOY, 1, "unbounded", [sigma, y, coeff])
        self.sigma, self.y, self.coeff = float(sigma), float(y), float(coeff)
        self.data = {"y_obs": np.array([self.y])}

    def forward_func(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        return theta[..., 1] ** 2 + self.coeff * theta[..., 0] ** 2 * (
            theta[..., 0] ** 2 - 0.5
        )

    def lift(self, theta):
        """theta -> q = (theta, eta) on the manifold (toy_2d_model.py:162-163)."""
        theta = np.asarray(theta, dtype=np.float64)
        eta = (self.y - self.forward_func(theta)) / self.sigma
        return np.concatenate([theta, eta[..., None]], -1)


class EightSchoolsModel(ModelSpec):
    """y = mu + tau v + sigma n with mu~N(0,5), tau~half-Cauchy(5).

    Reference: mlift/example_models/eight_schools.py:28-81.  Any number of "schools"
    compiled into the library (dim_y in {4, 8, 16}) is accepted.
    """

    family = "hierarchical"

    def __init__(self, y_obs, sigma, parametrization="unbounded"):
        y_obs = np.asarray(y_obs, dtype=np.float64)
        sigma = np.asarray(sigma, dtype=np.float64)
        super().__init__(_lib.MODEL_HIER, y_obs.size, parametrization,
                         np.concatenate([y_obs, sigma]))  # fmt: skip
        self.data = {"y_obs": y_obs, "σ": sigma, "parametrization": parametrization}

    def generate_params(self, u):
        """Host-side (NumPy) parameter transform, for initial states and traces."""
        u = np.asarray(u, dtype=np.float64)
        if self.parametrization == "unbounded":
            return {"μ": u[..., 0], "τ": np.exp(u[..., 1])}
        from scipy.special import ndtr

        return {"μ": 5.0 * u[..., 0],
                "τ": 5.0 * np.tan(np.pi * ((ndtr(u[..., 1]) + 1) / 2 - 0.5))}  # fmt: skip

    def sample_initial_states(self, rng, num_chain):
        """eight_schools.py:67-81: u ~ prior, v ~ N(0, I), n = (y - x) / sigma."""
        Y = self.dim_y
        if self.parametrization == "unbounded":
            u0 = rng.normal(0.0, 5.0, num_chain)
            u1 = np.log(np.abs(rng.standard_cauchy(num_chain) * 5.0))
        else:
            u0, u1 = rng.standard_normal(num_chain), rng.standard_normal(num_chain)
        u = np.stack([u0, u1], -1)
        v = rng.standard_normal((num_chain, Y))
        par = self.generate_params(u)
        x = par["μ"][:, None] + par["τ"][:, None] * v
        n = (self.data["y_obs"][None] - x) / self.data["σ"][None]
        return np.concatenate([u, v, n], -1)


def rk4_step_schedule(t_seq, dt):
    """Static RK4 step schedule of the reference (`_compute_dt_seq`, mlift/ode.py:45-53):
    between consecutive output times take floor(gap / dt) steps of `dt` plus one shorter
    step for the remainder.  Returns (dt_seq, index of the step landing on each output)."""
    t_seq = np.asarray(t_seq, dtype=np.float64)
    dt = float(np.asarray(dt).reshape(-1)[0])
    gaps = t_seq[1:] - t_seq[:-1]
    full_steps, remainders = np.divmod(gaps, dt)
    full_steps = full_steps.astype(np.int64)
    full_steps[remainders > 0] += 1
    last = full_steps.cumsum() - 1
    dt_seq = np.full(last[-1] + 1, dt)
    dt_seq[last[remainders > 0]] = remainders[remainders > 0]
    return dt_seq, last


class FitzHughNagumoModel(ModelSpec):
    """FitzHugh-Nagumo ODE model observed through x0 with additive noise.

    Reference: mlift/example_models/fitzhugh_nagumo.py:20-64 with the fixed-step RK4
    integrator of mlift/ode.py:9-53.  u = (alpha, beta, gamma, delta, epsilon, zeta, sigma
    [all log-normal priors, unconstrained through exp], x_init[2] ~ N(0, 1)); dim_u = 9,
    dim_y = len(y_obs), q = (u, n).  `data` keys as in the reference: y_obs, t_seq, dt.
    """

    family = "additive"
    LOC = np.array([0.0, 0.0, 0.0, -1.0, -3.0, -2.0, -1.0])  # fitzhugh_nagumo.py:20-28
    PARAM_NAMES = ("α", "β", "γ", "δ", "ϵ", "ζ", "σ")

    def __init__(self, y_obs, t_seq, dt, parametrization="unbounded"):
        y_obs = np.asarray(y_obs, dtype=np.float64).ravel()
        dt_seq, last = rk4_step_schedule(t_seq, dt)
        if last.size != y_obs.size:
            raise ValueError("t_seq must have len(y_obs) + 1 entries")
        blob = np.concatenate([[float(dt_seq.size)], y_obs, dt_seq, last.astype(np.float64)])
        super().__init__(_lib.MODEL_FN, y_obs.size, parametrization, blob)
        self.data = {"y_obs": y_obs, "t_seq": np.asarray(t_seq, dtype=np.float64),
                     "dt": float(np.asarray(dt).reshape(-1)[0]),
                     "parametrization": parametrization}  # fmt: skip
        self.dt_seq = dt_seq

    def generate_params(self, u):
        """Host-side parameter transform (prior.py:17-55), for initial states / traces."""
        u = np.asarray(u, dtype=np.float64)
        off = self.LOC if self.parametrization == "normal" else 0.0
        par = {k: np.exp(u[..., i] + (off[i] if self.parametrization == "normal" else 0.0))
               for i, k in enumerate(self.PARAM_NAMES)}  # fmt: skip
        par["x_init"] = u[..., 7:9]
        return par

    def lift(self, u):
        """On-manifold positions q = (u, n) with n = (y_obs - x(u)) / sigma(u)
        (fitzhugh_nagumo.py:118-120), the ODE solved on the GPU."""
        import ctypes as C

        import torch

        u = np.atleast_2d(np.asarray(u, dtype=np.float64))
        n = u.shape[0]
        q = _lib.to_soa(np.concatenate([u, np.zeros((n, self.dim_y))], 1))
        c = _lib.soa_empty(n, self.dim_y)
        pq, ld = _lib.p_soa(q, self.dim_q)
        _lib.check(_lib.lib().mlb_constr(self.handle, C.c_int64(n), C.c_int64(ld), pq,
                                         _lib.p_soa(c)[0], _lib.stream_ptr()))  # fmt: skip
        sigma = torch.as_tensor(self.generate_params(u)["σ"], device="cuda")
        q[:, self.dim_u :] = -c / sigma[:, None]  # c(u, n=0) = x(u) - y_obs
        return q


HH_CONSTANT_NAMES = (
    "k_alpha_n_1", "k_alpha_n_2", "k_alpha_n_3", "k_beta_n_1", "k_beta_n_2", "k_beta_n_3",
    "k_alpha_m_1", "k_alpha_m_2", "k_alpha_m_3", "k_beta_m_1", "k_beta_m_2", "k_beta_m_3",
    "k_alpha_h_1", "k_alpha_h_2", "k_alpha_h_3", "k_beta_h_1", "k_beta_h_2", "k_beta_h_3",
    "k_p_infinity_1", "k_p_infinity_2", "k_tau_p_2", "k_tau_p_3", "k_tau_p_4", "E_Na", "E_K",
    "E_leak", "c_m",
)  # fmt: skip


def hh_step_schedule(constants):
    """Time grid, per-step stimulus current and observation steps of the reference's
    `integrate_ode` / `generate_from_model` (hh_current_stimulus.py:117-122, 175-193):
    t = linspace(0, end_time, int(end_time / dt) + 1); the stimulus is on unless
    t < start or t > end; x_seq[start/dt : end/dt + 1 : interval/dt, 0] is observed."""
    k = {n: float(np.asarray(constants[n])) for n in
         ("end_time", "dt", "stimulus_start_time", "stimulus_end_time", "i_stimulus",
          "observation_interval")}  # fmt: skip
    n_point = int(k["end_time"] / k["dt"]) + 1
    t_seq = np.linspace(0, k["end_time"], n_point)
    t0, dt_seq = t_seq[:-1], t_seq[1:] - t_seq[:-1]
    stim = np.where((t0 < k["stimulus_start_time"]) | (t0 > k["stimulus_end_time"]), 0.0,
                    k["i_stimulus"])  # fmt: skip
    idx = np.arange(n_point)[slice(int(k["stimulus_start_time"] / k["dt"]),
                                   int(k["stimulus_end_time"] / k["dt"]) + 1,
                                   int(k["observation_interval"] / k["dt"]))]  # fmt: skip
    if idx[0] < 1:
        raise ValueError("observing the initial state is not supported")
    return dt_seq, stim, idx - 1  # x_seq[k] is the state after step k - 1


class HodgkinHuxleyModel(ModelSpec):
    """Hodgkin-Huxley current-stimulus model (5-state conductance model, exponential-Euler
    steps) observed through the membrane potential with additive noise.

    Reference: mlift/example_models/hh_current_stimulus.py:19-203.  u = (k_tau_p_1,
    g_bar_Na, g_bar_K, g_bar_M, g_leak, v_t, sigma); dim_u = 7, dim_y = len(y_obs) (1001
    for the reference's data file).  `constants`: the dict of the reference's
    `hodgkin-huxley-current-stimulus-data.npz` (rate constants, reversal potentials,
    stimulus timing, dt, observation_interval).
    """

    family = "additive"
    LOC = np.array([8.0, 4.0, 2.0, -3.0, -3.0, -60.0, 0.0])  # hh_current_stimulus.py:19-27
    SCALE = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 10.0, 1.0])
    PARAM_NAMES = ("k_tau_p_1", "g_bar_Na", "g_bar_K", "g_bar_M", "g_leak", "v_t", "σ")

    def __init__(self, y_obs, constants, parametrization="unbounded"):
        y_obs = np.asarray(y_obs, dtype=np.float64).ravel()
        dt_seq, stim, obs_step = hh_step_schedule(constants)
        if obs_step.size != y_obs.size:
            raise ValueError(f"y_obs must have {obs_step.size} entries for these constants")
        consts = np.array([float(np.asarray(constants[n])) for n in HH_CONSTANT_NAMES])
        blob = np.concatenate([[float(dt_seq.size)], y_obs, dt_seq,
                               obs_step.astype(np.float64), stim, consts])  # fmt: skip
        super().__init__(_lib.MODEL_HH, y_obs.size, parametrization, blob)
        self.data = {"y_obs": y_obs, "parametrization": parametrization,
                     **{n: float(np.asarray(constants[n])) for n in HH_CONSTANT_NAMES}}  # fmt: skip
        self.dt_seq, self.n_steps_used = dt_seq, int(obs_step[-1]) + 1

    def generate_params(self, u):
        u = np.asarray(u, dtype=np.float64)
        normal = self.parametrization == "normal"
        par = {}
        for i, k in enumerate(self.PARAM_NAMES):
            if k == "v_t":
                par[k] = self.LOC[i] + self.SCALE[i] * u[..., i] if normal else u[..., i]
            else:
                par[k] = np.exp(u[..., i] + (self.LOC[i] if normal else 0.0))
        return par

    def lift(self, u):
        """On-manifold positions q = (u, n), n = (y_obs - x(u)) / sigma(u), ODE on the GPU."""
        import ctypes as C

        import torch

        u = np.atleast_2d(np.asarray(u, dtype=np.float64))
        n = u.shape[0]
        q = _lib.to_soa(np.concatenate([u, np.zeros((n, self.dim_y))], 1))
        c = _lib.soa_empty(n, self.dim_y)
        pq, ld = _lib.p_soa(q, self.dim_q)
        _lib.check(_lib.lib().mlb_constr(self.handle, C.c_int64(n), C.c_int64(ld), pq,
                                         _lib.p_soa(c)[0], _lib.stream_ptr()))  # fmt: skip
        sigma = torch.as_tensor(self.generate_params(u)["σ"], device="cuda")
        q[:, self.dim_u :] = -c / sigma[:, None]
        return q
