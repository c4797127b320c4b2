// ODE-constrained models for the CPU oracle: FitzHugh-Nagumo, Hodgkin-Huxley.
// TEST INFRASTRUCTURE ONLY.
//
// Restates (paths under /root/reference/mlift/):
//   * example_models/fitzhugh_nagumo.py:20-58   priors, dx_dt, generate_y
//   * ode.py:9-42                               classic RK4 over a static step schedule
//     (the schedule itself, ode.py:45-53, is computed on the host by the caller and is
//     part of the model data: `dt_seq` and, per observation, the index of the step that
//     lands on it)
//   * prior.py:17-55, distributions.py:141-203  log-normal / normal priors under the
//     "unbounded" (x = exp(u)) and "normal" (x = exp(loc + scale u)) parametrisations
// The reference obtains the Jacobian by jax.jacfwd and the gradient of
// 0.5 log det(J J^T) by reverse-mode over that (systems.py:332-334, 568-580).  Here both
// come from a small forward-mode jet class (value, gradient, Hessian with respect to the
// ODE-related latent directions) pushed through the same RK4 recurrence - derived
// independently of the hand-written sensitivity equations in the CUDA kernels, and
// cross-checked against the torch autodiff restatement in oracle/mlift_oracle by
// tests/test_cpu_oracle_ode.py.
#include <array>

#include "oracle_models.hpp"

namespace orc {

// value + gradient (+ Hessian if H) with respect to N directions
template <int N, bool H>
struct Jet {
  static constexpr int NH = N * (N + 1) / 2;
  double v = 0.0;
  std::array<double, N> g{};
  std::array<double, H ? NH : 1> h{};
  Jet() {}
  Jet(double c) : v(c) {}
  static int idx(int a, int b) { return a <= b ? a * N - a * (a - 1) / 2 + (b - a)
                                               : b * N - b * (b - 1) / 2 + (a - b); }
};

template <int N, bool H>
Jet<N, H> operator+(const Jet<N, H>& x, const Jet<N, H>& y) {
  Jet<N, H> r;
  r.v = x.v + y.v;
  for (int a = 0; a < N; ++a) r.g[a] = x.g[a] + y.g[a];
  if (H)
    for (int k = 0; k < Jet<N, H>::NH; ++k) r.h[k] = x.h[k] + y.h[k];
  return r;
}
template <int N, bool H>
Jet<N, H> operator-(const Jet<N, H>& x, const Jet<N, H>& y) {
  Jet<N, H> r;
  r.v = x.v - y.v;
  for (int a = 0; a < N; ++a) r.g[a] = x.g[a] - y.g[a];
  if (H)
    for (int k = 0; k < Jet<N, H>::NH; ++k) r.h[k] = x.h[k] - y.h[k];
  return r;
}
template <int N, bool H>
Jet<N, H> operator-(const Jet<N, H>& x) {
  return Jet<N, H>(0.0) - x;
}
template <int N, bool H>
Jet<N, H> operator*(const Jet<N, H>& x, const Jet<N, H>& y) {
  Jet<N, H> r;
  r.v = x.v * y.v;
  for (int a = 0; a < N; ++a) r.g[a] = x.g[a] * y.v + x.v * y.g[a];
  if (H)
    for (int a = 0; a < N; ++a)
      for (int b = a; b < N; ++b) {
        int k = Jet<N, H>::idx(a, b);
        r.h[k] = x.h[k] * y.v + x.g[a] * y.g[b] + x.g[b] * y.g[a] + x.v * y.h[k];
      }
  return r;
}
template <int N, bool H>
Jet<N, H> operator*(double s, const Jet<N, H>& x) {
  Jet<N, H> r;
  r.v = s * x.v;
  for (int a = 0; a < N; ++a) r.g[a] = s * x.g[a];
  if (H)
    for (int k = 0; k < Jet<N, H>::NH; ++k) r.h[k] = s * x.h[k];
  return r;
}
template <int N, bool H>
Jet<N, H> operator/(const Jet<N, H>& x, double s) {
  Jet<N, H> r;
  r.v = x.v / s;
  for (int a = 0; a < N; ++a) r.g[a] = x.g[a] / s;
  if (H)
    for (int k = 0; k < Jet<N, H>::NH; ++k) r.h[k] = x.h[k] / s;
  return r;
}
// unary function with known first / second derivative at x.v
template <int N, bool H>
Jet<N, H> chain(const Jet<N, H>& x, double f, double df, double ddf) {
  Jet<N, H> r;
  r.v = f;
  for (int a = 0; a < N; ++a) r.g[a] = df * x.g[a];
  if (H)
    for (int a = 0; a < N; ++a)
      for (int b = a; b < N; ++b) {
        int k = Jet<N, H>::idx(a, b);
        r.h[k] = df * x.h[k] + ddf * x.g[a] * x.g[b];
      }
  return r;
}
template <int N, bool H>
Jet<N, H> jexp(const Jet<N, H>& x) {
  double e = std::exp(x.v);
  return chain(x, e, e, e);
}
template <int N, bool H>
Jet<N, H> jexpm1(const Jet<N, H>& x) {
  double e = std::expm1(x.v);
  return chain(x, e, e + 1.0, e + 1.0);
}
template <int N, bool H>
Jet<N, H> jrecip(const Jet<N, H>& x) {
  double r = 1.0 / x.v;
  return chain(x, r, -r * r, 2.0 * r * r * r);
}
template <int N, bool H>
Jet<N, H> operator/(const Jet<N, H>& x, const Jet<N, H>& y) { return x * jrecip(y); }
template <int N, bool H>
Jet<N, H> operator+(const Jet<N, H>& x, double c) { Jet<N, H> r = x; r.v += c; return r; }
template <int N, bool H>
Jet<N, H> operator+(double c, const Jet<N, H>& x) { return x + c; }
template <int N, bool H>
Jet<N, H> operator-(const Jet<N, H>& x, double c) { return x + (-c); }
template <int N, bool H>
Jet<N, H> operator-(double c, const Jet<N, H>& x) { return (-x) + c; }
template <int N, bool H>
Jet<N, H> operator*(const Jet<N, H>& x, double s) { return s * x; }
template <int N, bool H>
Jet<N, H> operator/(double c, const Jet<N, H>& x) { return c * jrecip(x); }
inline double jexp(double x) { return std::exp(x); }
inline double jexpm1(double x) { return std::expm1(x); }

// double as the order-0 "jet"
template <class T>
struct Lift {
  static T constant(double c) { return T(c); }
};

// ------------------------------------------------------------------ FitzHugh-Nagumo
struct FnModel : Model {
  static constexpr int ND = 8;  // jet directions: u0..u5 (ODE parameters), u7, u8 (x_init)
  int param, n_steps;
  std::vector<double> y_obs, dt_seq;
  std::vector<int> obs_step;
  double loc[7] = {0.0, 0.0, 0.0, -1.0, -3.0, -2.0, -1.0};  // fitzhugh_nagumo.py:20-28

  // data = {n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y]}
  FnModel(int Y_, int param_, const double* data) : param(param_) {
    family = FAMILY_ADDITIVE, U = 9, Y = Y_, Q = 9 + Y_, dense = false;
    n_steps = (int)data[0];
    y_obs.assign(data + 1, data + 1 + Y);
    dt_seq.assign(data + 1 + Y, data + 1 + Y + n_steps);
    for (int i = 0; i < Y; ++i) obs_step.push_back((int)data[1 + Y + n_steps + i]);
  }
  double offset(int k) const { return param == PARAM_NORMAL ? loc[k] : 0.0; }

  template <class T>
  static void dx_dt(const T (&x)[2], const T (&th)[6], T (&out)[2]) {
    // fitzhugh_nagumo.py:35-42
    out[0] = th[0] * x[0] - th[1] * (x[0] * x[0] * x[0]) + th[2] * x[1];
    out[1] = -(th[3] * x[0]) - th[4] * x[1] + th[5];
  }

  // RK4 over the schedule (ode.py:29-40); obs(i, x) is called when observation i is reached
  template <class T, class F>
  void integrate(const T (&th)[6], T (&x)[2], F&& obs) const {
    int next = 0;
    for (int s = 0; s < n_steps; ++s) {
      const double h = dt_seq[s];
      T k1[2], k2[2], k3[2], k4[2], xs[2];
      dx_dt(x, th, k1);
      for (int j = 0; j < 2; ++j) xs[j] = x[j] + (h * k1[j]) / 2.0;
      dx_dt(xs, th, k2);
      for (int j = 0; j < 2; ++j) xs[j] = x[j] + (h * k2[j]) / 2.0;
      dx_dt(xs, th, k3);
      for (int j = 0; j < 2; ++j) xs[j] = x[j] + h * k3[j];
      dx_dt(xs, th, k4);
      for (int j = 0; j < 2; ++j)
        x[j] = x[j] + (h * (k1[j] + 2.0 * k2[j] + 2.0 * k3[j] + k4[j])) / 6.0;
      while (next < Y && obs_step[next] == s) obs(next++, x);
    }
  }

  // seeds: direction d of the jet <-> latent index dir_u(d)
  static int dir_u(int d) { return d < 6 ? d : d + 1; }

  template <class T>
  void seed(const double* u, T (&th)[6], T (&x)[2]) const {
    for (int k = 0; k < 6; ++k) {
      T uk(u[k] + offset(k));
      seed_dir(uk, k);
      th[k] = jexp(uk);
    }
    for (int j = 0; j < 2; ++j) {
      x[j] = T(u[7 + j]);
      seed_dir(x[j], 6 + j);
    }
  }
  static void seed_dir(double&, int) {}
  template <int N, bool H>
  static void seed_dir(Jet<N, H>& x, int d) { x.g[d] = 1.0; }

  void constr(const double* q, double* c) const override {
    double th[6], x[2];
    seed<double>(q, th, x);
    const double sigma = std::exp(q[6] + offset(6));
    const double* n = q + 9;
    integrate<double>(th, x, [&](int i, const double (&xx)[2]) {
      c[i] = xx[0] + sigma * n[i] - y_obs[i];
    });
  }
  void jacob(const double* q, double* dy_du, double*, double* dy_dn,
             double* c) const override {
    using J1 = Jet<ND, false>;
    J1 th[6], x[2];
    seed<J1>(q, th, x);
    const double sigma = std::exp(q[6] + offset(6));
    const double* n = q + 9;
    integrate<J1>(th, x, [&](int i, const J1 (&xx)[2]) {
      c[i] = xx[0].v + sigma * n[i] - y_obs[i];
      for (int d = 0; d < ND; ++d) dy_du[(size_t)i * 9 + dir_u(d)] = xx[0].g[d];
      dy_du[(size_t)i * 9 + 6] = sigma * n[i];  // d(sigma n_i)/du6, sigma = exp(u6 + off)
      dy_dn[i] = sigma;
    });
  }
  double nld(const double* q, double* grad) const override {
    // log-normal priors pulled back through x = exp(u): ((u - loc)/scale)^2 / 2 (the
    // log x terms of distributions.py:182-183 and :78-80 cancel); "normal": u^2 / 2
    double val = 0.0;
    for (int k = 0; k < 7; ++k) {
      double z = q[k] - (param == PARAM_NORMAL ? 0.0 : loc[k]);
      val += z * z / 2.0;
      if (grad) grad[k] = z;
    }
    for (int k = 7; k < Q; ++k) {
      val += q[k] * q[k] / 2.0;
      if (grad) grad[k] = q[k];
    }
    return val;
  }
  // d/dq of 0.5 log det(J J^T) = 0.5 log det C + Y log sigma,  C = I + A^T A / sigma^2.
  // With M = C^-1 and B = A M / sigma^2:
  //   g_u[b]  = sum_i sum_a B_ia d2x_i/du_a du_b                   (b an ODE direction)
  //   g_u[6]  = Y - (U - tr M) + sum_i B_i6 sigma n_i
  //   g_n[i]  = B_i6 sigma
  void grad_logdet(const double* q, const double* dy_du, const double*, const double* dy_dn,
                   const double* chol, const double*, double* g) const override {
    const int Un = 9;
    const double sigma = dy_dn[0];
    // M = C^-1 from the lower Cholesky factor, column by column
    double Minv[81];
    for (int j = 0; j < Un; ++j) {
      double col[9];
      for (int i = 0; i < Un; ++i) col[i] = (i == j) ? 1.0 : 0.0;
      for (int i = 0; i < Un; ++i) {
        double t = col[i];
        for (int k = 0; k < i; ++k) t -= chol[i * Un + k] * col[k];
        col[i] = t / chol[i * Un + i];
      }
      for (int i = Un - 1; i >= 0; --i) {
        double t = col[i];
        for (int k = i + 1; k < Un; ++k) t -= chol[k * Un + i] * col[k];
        col[i] = t / chol[i * Un + i];
      }
      for (int i = 0; i < Un; ++i) Minv[i * Un + j] = col[i];
    }
    std::vector<double> B((size_t)Y * Un);
    for (int i = 0; i < Y; ++i)
      for (int a = 0; a < Un; ++a) {
        double s = 0.0;
        for (int k = 0; k < Un; ++k) s += dy_du[(size_t)i * Un + k] * Minv[k * Un + a];
        B[(size_t)i * Un + a] = s / (sigma * sigma);
      }
    for (int k = 0; k < Q; ++k) g[k] = 0.0;
    double trM = 0.0;
    for (int a = 0; a < Un; ++a) trM += Minv[a * Un + a];
    const double* n = q + 9;
    double g6 = (double)Y - ((double)Un - trM);
    for (int i = 0; i < Y; ++i) {
      g6 += B[(size_t)i * Un + 6] * sigma * n[i];
      g[9 + i] = B[(size_t)i * Un + 6] * sigma;
    }
    g[6] = g6;
    using J2 = Jet<ND, true>;
    J2 th[6], x[2];
    seed<J2>(q, th, x);
    integrate<J2>(th, x, [&](int i, const J2 (&xx)[2]) {
      for (int b = 0; b < ND; ++b) {
        double s = 0.0;
        for (int a = 0; a < ND; ++a)
          s += B[(size_t)i * Un + dir_u(a)] * xx[0].h[J2::idx(a, b)];
        g[dir_u(b)] += s;
      }
    });
  }
};


// ------------------------------------------------------------------ Hodgkin-Huxley
// hh_current_stimulus.py:19-203: 5-state conductance model (v, n, m, h, p) advanced with
// exponential-Euler steps over t = linspace(0, end_time, n) (`integrate_ode`, :142-178),
// initialised at the steady state for v = E_leak (:125-139), observed through v at the
// stimulus window (:189-193).  u = (k_tau_p_1, g_bar_Na, g_bar_K, g_bar_M, g_leak, v_t,
// sigma): five log-normal parameters, v_t ~ N(-60, 10), sigma log-normal (:19-27).
// data = {n_steps, y_obs[Y], dt_seq[n_steps], obs_step[Y], stim[n_steps], consts[27]}
// with consts in the order of HH_CONSTS below (the per-step stimulus current, :117-122, is
// evaluated on the host from the same comparisons).
struct HhModel : Model {
  static constexpr int ND = 6, NC = 27;
  enum { k_alpha_n_1, k_alpha_n_2, k_alpha_n_3, k_beta_n_1, k_beta_n_2, k_beta_n_3,
         k_alpha_m_1, k_alpha_m_2, k_alpha_m_3, k_beta_m_1, k_beta_m_2, k_beta_m_3,
         k_alpha_h_1, k_alpha_h_2, k_alpha_h_3, k_beta_h_1, k_beta_h_2, k_beta_h_3,
         k_p_infinity_1, k_p_infinity_2, k_tau_p_2, k_tau_p_3, k_tau_p_4, E_Na, E_K, E_leak,
         c_m };
  int param, n_steps;
  std::vector<double> y_obs, dt_seq, stim, K;
  std::vector<int> obs_step;
  double loc[7] = {8.0, 4.0, 2.0, -3.0, -3.0, -60.0, 0.0};
  double scale[7] = {1.0, 1.0, 1.0, 1.0, 1.0, 10.0, 1.0};

  HhModel(int Y_, int param_, const double* d) : param(param_) {
    family = FAMILY_ADDITIVE, U = 7, Y = Y_, Q = 7 + Y_, dense = false;
    n_steps = (int)d[0];
    const double* p = d + 1;
    y_obs.assign(p, p + Y), p += Y;
    dt_seq.assign(p, p + n_steps), p += n_steps;
    for (int i = 0; i < Y; ++i) obs_step.push_back((int)p[i]);
    p += Y;
    stim.assign(p, p + n_steps), p += n_steps;
    K.assign(p, p + NC);
  }

  template <class T> static T xoe(const T& x) { return x / jexpm1(x); }
  static double xoe(double x) { return x / std::expm1(x); }

  template <class T>
  struct Par {
    T k_tau_p_1, g_Na, g_K, g_M, g_leak, v_t;
  };
  template <class T> T alpha_n(const T& v, const Par<T>& p) const {
    return K[k_alpha_n_1] * xoe(-(v - p.v_t - K[k_alpha_n_2]) / K[k_alpha_n_3]) * K[k_alpha_n_3];
  }
  template <class T> T beta_n(const T& v, const Par<T>& p) const {
    return K[k_beta_n_1] * jexp(-(v - p.v_t - K[k_beta_n_2]) / K[k_beta_n_3]);
  }
  template <class T> T alpha_m(const T& v, const Par<T>& p) const {
    return K[k_alpha_m_1] * xoe(-(v - p.v_t - K[k_alpha_m_2]) / K[k_alpha_m_3]) * K[k_alpha_m_3];
  }
  template <class T> T beta_m(const T& v, const Par<T>& p) const {
    return K[k_beta_m_1] * xoe((v - p.v_t - K[k_beta_m_2]) / K[k_beta_m_3]) * K[k_beta_m_3];
  }
  template <class T> T alpha_h(const T& v, const Par<T>& p) const {
    return K[k_alpha_h_1] * jexp(-(v - p.v_t - K[k_alpha_h_2]) / K[k_alpha_h_3]);
  }
  template <class T> T beta_h(const T& v, const Par<T>& p) const {
    return K[k_beta_h_1] / (jexp(-(v - p.v_t - K[k_beta_h_2]) / K[k_beta_h_3]) + 1.0);
  }
  template <class T> T p_inf(const T& v) const {
    return 1.0 / (1.0 + jexp((-v + K[k_p_infinity_1]) / K[k_p_infinity_2]));
  }
  template <class T> T tau_p(const T& v, const Par<T>& p) const {
    return p.k_tau_p_1 / (K[k_tau_p_2] * jexp((v + K[k_tau_p_3]) / K[k_tau_p_4]) +
                          jexp(-(v + K[k_tau_p_3]) / K[k_tau_p_4]));
  }

  static void seed_dir(double&, int) {}
  template <int N, bool H>
  static void seed_dir(Jet<N, H>& x, int d) { x.g[d] = 1.0; }

  template <class T>
  Par<T> params(const double* u) const {
    T t[6];
    for (int k = 0; k < 6; ++k) {
      T uk(u[k]);
      seed_dir(uk, k);
      if (k < 5)
        t[k] = jexp(param == PARAM_NORMAL ? uk + loc[k] : uk);
      else
        t[k] = param == PARAM_NORMAL ? loc[k] + scale[k] * uk : uk;
    }
    return Par<T>{t[0], t[1], t[2], t[3], t[4], t[5]};
  }
  double sigma(const double* u) const {
    return std::exp(u[6] + (param == PARAM_NORMAL ? loc[6] : 0.0));
  }

  template <class T, class F>
  void integrate(const Par<T>& p, F&& obs) const {
    T v(K[E_leak]);
    T an = alpha_n(v, p), bn = beta_n(v, p), am = alpha_m(v, p), bm = beta_m(v, p);
    T ah = alpha_h(v, p), bh = beta_h(v, p);
    T n = an / (an + bn), m = am / (am + bm), h = ah / (ah + bh), pp = p_inf(v);
    int next = 0;
    for (int s = 0; s < n_steps && next < Y; ++s) {
      const double dt = dt_seq[s];
      T g_K = p.g_K * ((n * n) * (n * n));
      T g_Na = p.g_Na * (m * m * m) * h;
      T g_M = p.g_M * pp;
      T tau_v = K[c_m] / (g_K + g_Na + g_M + p.g_leak);
      T v_inf = (g_K * K[E_K] + g_Na * K[E_Na] + g_M * K[E_K] + p.g_leak * K[E_leak] + stim[s]) *
                tau_v / K[c_m];
      T v_ = v_inf + (v - v_inf) * jexp(-(dt / tau_v));
      an = alpha_n(v_, p), bn = beta_n(v_, p), am = alpha_m(v_, p), bm = beta_m(v_, p);
      ah = alpha_h(v_, p), bh = beta_h(v_, p);
      T tau_n = 1.0 / (an + bn), n_inf = an / (an + bn);
      T tau_m = 1.0 / (am + bm), m_inf = am / (am + bm);
      T tau_h = 1.0 / (ah + bh), h_inf = ah / (ah + bh);
      n = n_inf + (n - n_inf) * jexp(-(dt / tau_n));
      m = m_inf + (m - m_inf) * jexp(-(dt / tau_m));
      h = h_inf + (h - h_inf) * jexp(-(dt / tau_h));
      T pi = p_inf(v_);
      pp = pi + (pp - pi) * jexp(-(dt / tau_p(v_, p)));
      v = v_;
      while (next < Y && obs_step[next] == s) obs(next++, v);
    }
  }

  void constr(const double* q, double* c) const override {
    const double sg = sigma(q);
    const double* n = q + 7;
    integrate<double>(params<double>(q), [&](int i, const double& v) {
      c[i] = v + sg * n[i] - y_obs[i];
    });
  }
  void jacob(const double* q, double* dy_du, double*, double* dy_dn,
             double* c) const override {
    using J1 = Jet<ND, false>;
    const double sg = sigma(q);
    const double* n = q + 7;
    integrate<J1>(params<J1>(q), [&](int i, const J1& v) {
      c[i] = v.v + sg * n[i] - y_obs[i];
      for (int d = 0; d < ND; ++d) dy_du[(size_t)i * 7 + d] = v.g[d];
      dy_du[(size_t)i * 7 + 6] = sg * n[i];
      dy_dn[i] = sg;
    });
  }
  double nld(const double* q, double* grad) const override {
    double val = 0.0;
    for (int k = 0; k < 7; ++k) {
      double z = param == PARAM_NORMAL ? q[k] : (q[k] - loc[k]) / scale[k];
      val += z * z / 2.0;
      if (grad) grad[k] = param == PARAM_NORMAL ? z : z / scale[k];
    }
    for (int k = 7; k < Q; ++k) {
      val += q[k] * q[k] / 2.0;
      if (grad) grad[k] = q[k];
    }
    return val;
  }
  void grad_logdet(const double* q, const double* dy_du, const double*, const double* dy_dn,
                   const double* chol, const double*, double* g) const override {
    const int Un = 7;
    const double sg = dy_dn[0];
    double Minv[49];
    for (int j = 0; j < Un; ++j) {
      double col[7];
      for (int i = 0; i < Un; ++i) col[i] = (i == j) ? 1.0 : 0.0;
      for (int i = 0; i < Un; ++i) {
        double t = col[i];
        for (int k = 0; k < i; ++k) t -= chol[i * Un + k] * col[k];
        col[i] = t / chol[i * Un + i];
      }
      for (int i = Un - 1; i >= 0; --i) {
        double t = col[i];
        for (int k = i + 1; k < Un; ++k) t -= chol[k * Un + i] * col[k];
        col[i] = t / chol[i * Un + i];
      }
      for (int i = 0; i < Un; ++i) Minv[i * Un + j] = col[i];
    }
    std::vector<double> B((size_t)Y * Un);
    for (int i = 0; i < Y; ++i)
      for (int a = 0; a < Un; ++a) {
        double s = 0.0;
        for (int k = 0; k < Un; ++k) s += dy_du[(size_t)i * Un + k] * Minv[k * Un + a];
        B[(size_t)i * Un + a] = s / (sg * sg);
      }
    for (int k = 0; k < Q; ++k) g[k] = 0.0;
    double trM = 0.0;
    for (int a = 0; a < Un; ++a) trM += Minv[a * Un + a];
    const double* n = q + 7;
    double g6 = (double)Y - ((double)Un - trM);
    for (int i = 0; i < Y; ++i) {
      g6 += B[(size_t)i * Un + 6] * sg * n[i];
      g[7 + i] = B[(size_t)i * Un + 6] * sg;
    }
    g[6] = g6;
    using J2 = Jet<ND, true>;
    integrate<J2>(params<J2>(q), [&](int i, const J2& v) {
      for (int b = 0; b < ND; ++b) {
        double s = 0.0;
        for (int a = 0; a < ND; ++a) s += B[(size_t)i * Un + a] * v.h[J2::idx(a, b)];
        g[b] += s;
      }
    });
  }
};

Model* make_hh_model(int dim_y, int param, const double* data, int n_data) {
  if (n_data < 1) return nullptr;
  int n_steps = (int)data[0];
  if (n_steps <= 0 || n_data < 1 + 2 * dim_y + 2 * n_steps + HhModel::NC) return nullptr;
  return new HhModel(dim_y, param, data);
}

Model* make_fn_model(int dim_y, int param, const double* data, int n_data) {
  if (n_data < 1 + dim_y) return nullptr;
  int n_steps = (int)data[0];
  if (n_steps <= 0 || n_data < 1 + 2 * dim_y + n_steps) return nullptr;
  return new FnModel(dim_y, param, data);
}
}  // namespace orc
