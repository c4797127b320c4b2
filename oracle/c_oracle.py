"""ctypes binding of the closed-form C++ CPU oracle (oracle/c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
Arrays are NumPy fp64, chains as rows ([n_chain, dim], C-contiguous).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

MODEL_TOY, MODEL_HIER, MODEL_FN, MODEL_HH = 0, 1, 2, 3
PARAM_UNBOUNDED, PARAM_NORMAL = 0, 1
(SOLVER_QUASI_NEWTON, SOLVER_NEWTON, SOLVER_NEWTON_LS, SOLVER_MICI_NEWTON,
 SOLVER_MICI_QUASI_NEWTON) = range(5)  # fmt: skip
ST_OK, ST_DIVERGED, ST_NOT_CONVERGED, ST_NON_REVERSIBLE, ST_H_DIVERGED, ST_LINALG = range(6)
NORM_MAX, NORM_EUCLID = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "c")
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", src] + (["-B"] if force else []))
    return LIB_PATH


class SolverOpts(C.Structure):
    _fields_ = [
        ("solver", C.c_int), ("max_iters", C.c_int), ("max_ls", C.c_int),
        ("norm", C.c_int), ("ctol", C.c_double), ("ptol", C.c_double),
        ("dtol", C.c_double), ("rev_tol", C.c_double), ("rev_norm", C.c_int),
        ("flags", C.c_int),  # NUTS_* (dynamic sampler only)
    ]  # fmt: skip


def solver_opts(solver=SOLVER_NEWTON, constraint_tol=1e-9, position_tol=1e-8,
                divergence_tol=1e10, max_iters=50, max_line_search_iters=10,
                norm=NORM_MAX, reverse_check_tol=None, reverse_check_norm=NORM_MAX,
                flags=0):  # fmt: skip
    if reverse_check_tol is None:
        reverse_check_tol = 2 * position_tol
    return SolverOpts(solver, max_iters, max_line_search_iters, norm, constraint_tol,
                      position_tol, divergence_tol, reverse_check_tol,
                      reverse_check_norm, flags)  # fmt: skip


# orc_solver_opts.flags (same values as include/mlb.h MLB_FLAG_NUTS_*): the dynamic sampler
# follows Mici's defaults (extra subtree checks, one-sided h - h_init > max_delta_h)
NUTS_NO_EXTRA_SUBTREE_CHECKS, NUTS_TWO_SIDED_DELTA_H = 2, 4


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_model_create.restype = C.c_void_p
        _lib.orc_model_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        _lib.orc_philox_uniform.restype = C.c_double
        _lib.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib.orc_philox_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int,
                                            C.c_void_p]  # fmt: skip
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleModel:
    def __init__(self, model_id, dim_y, param, data):
        data = _f(data).ravel()
        self.h = C.c_void_p(lib().orc_model_create(model_id, dim_y, param, _p(data),
                                                   data.size))  # fmt: skip
        if not self.h:
            raise ValueError("oracle model could not be created")
        L = lib()
        self.U, self.Y, self.Q = (L.orc_dim_u(self.h), L.orc_dim_y(self.h),
                                  L.orc_dim_q(self.h))  # fmt: skip
        self.dense = bool(L.orc_is_dense(self.h))
        self.K = self.Y if self.dense else self.U
        self.hier = model_id == MODEL_HIER
        self.model_id = model_id

    @classmethod
    def toy(cls, sigma, y=1.0, coeff=2.0):
        return cls(MODEL_TOY, 1, 0, [sigma, y, coeff])

    @classmethod
    def hier(cls, y_obs, sigma, parametrization="unbounded"):
        y_obs, sigma = _f(y_obs), _f(sigma)
        par = PARAM_NORMAL if parametrization == "normal" else PARAM_UNBOUNDED
        return cls(MODEL_HIER, y_obs.size, par, np.concatenate([y_obs, sigma]))

    @classmethod
    def fn(cls, y_obs, t_seq, dt, parametrization="unbounded"):
        """FitzHugh-Nagumo: the static RK4 step schedule (reference ode.py:45-53) is
        computed here on the host and passed as data."""
        data, _ = fn_model_data(y_obs, t_seq, dt)
        par = PARAM_NORMAL if parametrization == "normal" else PARAM_UNBOUNDED
        return cls(MODEL_FN, int(np.asarray(y_obs).size), par, data)

    @classmethod
    def hh(cls, y_obs, constants, parametrization="unbounded"):
        """Hodgkin-Huxley current-stimulus model; `constants` is the dict of the
        reference's data file (hh_current_stimulus.py:287-292)."""
        data = hh_model_data(y_obs, constants)
        par = PARAM_NORMAL if parametrization == "normal" else PARAM_UNBOUNDED
        return cls(MODEL_HH, int(np.asarray(y_obs).size), par, data)

    def __del__(self):
        try:
            lib().orc_model_destroy(self.h)
        except Exception:
            pass

    # ---- individual operations (one per row of SURVEY 2b)
    def constr(self, q):
        q = _f(q)
        c = np.empty((q.shape[0], self.Y))
        lib().orc_constr(self.h, C.c_int64(q.shape[0]), _p(q), _p(c))
        return c

    def jacob_constr_blocks(self, q):
        q = _f(q)
        n = q.shape[0]
        dy_du = np.empty((n, self.Y, self.U))
        dy_dv = np.empty((n, self.Y)) if self.hier else None
        dy_dn, c = np.empty((n, self.Y)), np.empty((n, self.Y))
        lib().orc_jacob_constr_blocks(self.h, C.c_int64(n), _p(q), _p(dy_du), _p(dy_dv),
                                      _p(dy_dn), _p(c))  # fmt: skip
        return (dy_du, dy_dv, dy_dn), c

    def decompose_gram(self, jac):
        dy_du, dy_dv, dy_dn = jac
        n = dy_du.shape[0]
        chol, d = np.empty((n, self.K, self.K)), np.empty((n, self.Y))
        lib().orc_decompose_gram(self.h, C.c_int64(n), _p(dy_du), _p(dy_dv), _p(dy_dn),
                                 _p(chol), _p(d))  # fmt: skip
        return chol, d

    def log_det_sqrt_gram(self, q):
        q = _f(q)
        val = np.empty(q.shape[0])
        lib().orc_log_det_sqrt_gram(self.h, C.c_int64(q.shape[0]), _p(q), _p(val))
        return val

    def grad_log_det_sqrt_gram(self, q):
        q = _f(q)
        val, g = np.empty(q.shape[0]), np.empty_like(q)
        lib().orc_grad_log_det_sqrt_gram(self.h, C.c_int64(q.shape[0]), _p(q), _p(g),
                                         _p(val))  # fmt: skip
        return g, val

    def neg_log_dens(self, q, with_grad=True):
        q = _f(q)
        val = np.empty(q.shape[0])
        g = np.empty_like(q) if with_grad else None
        lib().orc_neg_log_dens(self.h, C.c_int64(q.shape[0]), _p(q), _p(val), _p(g))
        return (val, g) if with_grad else val

    def lmult_by_jacob_constr(self, jac, v):
        v = _f(v)
        out = np.empty((v.shape[0], self.Y))
        lib().orc_lmult_by_jacob_constr(self.h, C.c_int64(v.shape[0]), _p(jac[0]),
                                        _p(jac[1]), _p(jac[2]), _p(v), _p(out))  # fmt: skip
        return out

    def rmult_by_jacob_constr(self, jac, w):
        w = _f(w)
        out = np.empty((w.shape[0], self.Q))
        lib().orc_rmult_by_jacob_constr(self.h, C.c_int64(w.shape[0]), _p(jac[0]),
                                        _p(jac[1]), _p(jac[2]), _p(w), _p(out))  # fmt: skip
        return out

    def lmult_by_pinv_jacob_constr(self, jac, gram, w):
        w = _f(w)
        out = np.empty((w.shape[0], self.Q))
        lib().orc_lmult_by_pinv_jacob_constr(
            self.h, C.c_int64(w.shape[0]), _p(jac[0]), _p(jac[1]), _p(jac[2]),
            _p(gram[0]), _p(gram[1]), _p(w), _p(out))  # fmt: skip
        return out

    def normal_space_component(self, jac, gram, v):
        v = _f(v)
        out = np.empty((v.shape[0], self.Q))
        lib().orc_normal_space_component(
            self.h, C.c_int64(v.shape[0]), _p(jac[0]), _p(jac[1]), _p(jac[2]),
            _p(gram[0]), _p(gram[1]), _p(v), _p(out))  # fmt: skip
        return out

    def lmult_by_inv_jacob_product(self, jac_l, jac_r, w):
        w = _f(w)
        out = np.empty((w.shape[0], self.Y))
        lib().orc_lmult_by_inv_jacob_product(
            self.h, C.c_int64(w.shape[0]), _p(jac_l[0]), _p(jac_l[1]), _p(jac_l[2]),
            _p(jac_r[0]), _p(jac_r[1]), _p(jac_r[2]), _p(w), _p(out))  # fmt: skip
        return out

    @staticmethod
    def last_line_search_ambiguous(n):
        """[n] bool: chains of the last solve_projection / leapfrog_steps call in which a
        backtracking comparison of the line search (systems.py:299-303) was decided by
        rounding (both residual norms below 1e-10, or equal to 1e-6 relative)."""
        out = np.zeros(n, np.int32)
        lib().orc_last_line_search_ambiguous(_p(out), C.c_int64(n))
        return out.astype(bool)

    # ---- projection / step / transition drivers
    def solve_projection(self, q, q_prev, dt, opts):
        q, q_prev = _f(q), _f(q_prev)
        n = q.shape[0]
        q_out, mu = np.empty_like(q), np.empty_like(q)
        iters, n_constr, status = (np.empty(n, np.int32) for _ in range(3))
        ndq, err = np.empty(n), np.empty(n)
        lib().orc_solve_projection(
            self.h, C.c_int64(n), _p(q), _p(q_prev), C.c_double(dt), C.byref(opts),
            _p(q_out), _p(mu), _p(iters), _p(ndq), _p(err), _p(n_constr), _p(status))  # fmt: skip
        return dict(q=q_out, mu=mu, iters=iters, norm_dq=ndq, err=err,
                    n_constr=n_constr, status=status)  # fmt: skip

    def leapfrog_steps(self, q, p, dt, n_step, opts, dt_per_chain=None):
        q, p = _f(q).copy(), _f(p).copy()
        n = q.shape[0]
        status, n_done, itf, itr = (np.empty(n, np.int32) for _ in range(4))
        h0, h1 = np.empty(n), np.empty(n)
        dpc = None if dt_per_chain is None else _f(dt_per_chain)
        lib().orc_leapfrog_steps(
            self.h, C.c_int64(n), _p(q), _p(p), C.c_double(dt), _p(dpc), C.c_int(n_step),
            C.byref(opts), _p(status), _p(n_done), _p(itf), _p(itr), _p(h0), _p(h1))  # fmt: skip
        return dict(q=q, p=p, status=status, n_done=n_done, iters_fwd=itf, iters_rev=itr,
                    h_init=h0, h_final=h1)  # fmt: skip

    def nuts_transition(self, q, dt, opts, max_depth=10, mom_noise=None, unif=None, seed=0,
                        chain0=0, iteration=0, max_delta_h=1000.0, dt_per_chain=None):  # fmt: skip
        """One dynamic multinomial HMC transition per chain (oracle/c/chmc_oracle.cpp
        `orc_nuts_transition`).  unif: optional [n, n_unif] supplied uniforms."""
        q = _f(q).copy()
        n = q.shape[0]
        status, n_step, depth = (np.empty(n, np.int32) for _ in range(3))
        a_stat = np.empty(n)
        mom_noise = None if mom_noise is None else _f(mom_noise)
        unif = None if unif is None else _f(unif)
        dpc = None if dt_per_chain is None else _f(dt_per_chain)
        lib().orc_nuts_transition(
            self.h, C.c_int64(n), _p(q), C.c_double(dt), _p(dpc), C.c_int(max_depth),
            C.byref(opts), C.c_double(max_delta_h), _p(mom_noise), _p(unif),
            C.c_int(0 if unif is None else unif.shape[1]), C.c_uint64(seed),
            C.c_uint64(chain0), C.c_uint32(iteration), _p(status), _p(a_stat), _p(n_step),
            _p(depth))  # fmt: skip
        return dict(q=q, status=status, accept_stat=a_stat, n_step=n_step, depth=depth)

    def hmc_transition(self, q, dt, n_step, opts, mom_noise=None, unif=None, seed=0,
                       chain0=0, iteration=0, max_delta_h=1000.0, dt_per_chain=None):  # fmt: skip
        q = _f(q).copy()
        n = q.shape[0]
        status, acc, n_done, its = (np.empty(n, np.int32) for _ in range(4))
        a_stat, h = np.empty(n), np.empty(n)
        mom_noise = None if mom_noise is None else _f(mom_noise)
        unif = None if unif is None else _f(unif)
        dpc = None if dt_per_chain is None else _f(dt_per_chain)
        lib().orc_hmc_transition(
            self.h, C.c_int64(n), _p(q), C.c_double(dt), _p(dpc), C.c_int(n_step),
            C.byref(opts), C.c_double(max_delta_h), _p(mom_noise), _p(unif),
            C.c_uint64(seed), C.c_uint64(chain0), C.c_uint32(iteration), _p(status),
            _p(acc), _p(a_stat), _p(n_done), _p(its), _p(h))  # fmt: skip
        return dict(q=q, status=status, accepted=acc, accept_stat=a_stat, n_done=n_done,
                    iters_total=its, h=h)  # fmt: skip


def rk4_schedule(t_seq, dt):
    """Step sizes and, per output time, the index of the step landing on it (reference
    mlift/ode.py:45-53 `_compute_dt_seq`)."""
    t_seq = np.asarray(t_seq, dtype=np.float64)
    dt = float(np.asarray(dt).reshape(-1)[0])
    gaps = t_seq[1:] - t_seq[:-1]
    n_full, rem = np.divmod(gaps, dt)
    n_full = n_full.astype(np.int64)
    n_full[rem > 0] += 1
    last = n_full.cumsum() - 1
    dt_seq = np.full(last[-1] + 1, dt)
    dt_seq[last[rem > 0]] = rem[rem > 0]
    return dt_seq, last


def fn_model_data(y_obs, t_seq, dt):
    """Data blob of the FN model: {n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y]}."""
    dt_seq, last = rk4_schedule(t_seq, dt)
    y_obs = _f(y_obs).ravel()
    assert last.size == y_obs.size
    return np.concatenate([[float(dt_seq.size)], y_obs, dt_seq, last.astype(np.float64)]), dt_seq


HH_CONSTS = (
    "k_alpha_n_1", "k_alpha_n_2", "k_alpha_n_3", "k_beta_n_1", "k_beta_n_2", "k_beta_n_3",
    "k_alpha_m_1", "k_alpha_m_2", "k_alpha_m_3", "k_beta_m_1", "k_beta_m_2", "k_beta_m_3",
    "k_alpha_h_1", "k_alpha_h_2", "k_alpha_h_3", "k_beta_h_1", "k_beta_h_2", "k_beta_h_3",
    "k_p_infinity_1", "k_p_infinity_2", "k_tau_p_2", "k_tau_p_3", "k_tau_p_4", "E_Na", "E_K",
    "E_leak", "c_m",
)  # fmt: skip


def hh_schedule(constants):
    """Time grid, per-step stimulus current and observation steps of the reference's
    `integrate_ode` / `generate_from_model` (hh_current_stimulus.py:117-122, 175-193)."""
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
    assert idx[0] >= 1, "observing the initial state is not supported"
    return dt_seq, stim, idx - 1  # x_seq[k] is the state after step k - 1


def hh_model_data(y_obs, constants):
    """{n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y], stim[n_steps], consts[27]}."""
    dt_seq, stim, obs_step = hh_schedule(constants)
    y_obs = _f(y_obs).ravel()
    assert obs_step.size == y_obs.size
    consts = np.array([float(np.asarray(constants[n])) for n in HH_CONSTS])
    return np.concatenate([[float(dt_seq.size)], y_obs, dt_seq, obs_step.astype(np.float64),
                           stim, consts])  # fmt: skip


def philox_normals(seed, chain, iteration, n):
    out = np.empty(n + (n & 1))
    lib().orc_philox_normals(seed, chain, iteration, n, _p(out))
    return out[:n]


def philox_uniform(seed, chain, iteration):
    return lib().orc_philox_uniform(seed, chain, iteration)


def max_threads():
    return lib().orc_max_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says
    (torchrun sets it to 1 for its workers).  Returns the count."""
    import os

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(C.c_int(n))
    return max_threads()
