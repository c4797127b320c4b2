// FitzHugh-Nagumo ODE-constrained model: RK4 sweeps with hand-derived first- and
// second-order forward sensitivities, one chain per thread, ODE state in registers.
//
// Reference: mlift/example_models/fitzhugh_nagumo.py:20-58 (priors, dx_dt, generate_y),
// mlift/ode.py:9-42 (classic RK4 over a static step schedule; the schedule of ode.py:45-53
// is computed on the host and is model data).  The reference differentiates the RK4
// recurrence with jax.jacfwd (Jacobian, systems.py:568-580) and reverse-over-forward
// (gradient of 0.5 log det J J^T, systems.py:332-334); here the same derivatives of the
// same discrete recurrence come from integrating the sensitivity recurrences alongside
// the state (forward-over-forward), which needs no tape.
//
// Latent layout: u = (log-ish alpha, beta, gamma, delta, epsilon, zeta, sigma, x_init[2]),
// theta_k = exp(u_k + off_k) with off_k = prior location under the "normal"
// parametrisation (prior.py:45-55, distributions.py:190-195) and 0 under "unbounded"
// (transforms.py:43-55).  Sensitivity directions d = 0..5 <-> u0..u5, d = 6,7 <-> u7,u8.
#pragma once
#include <utility>

#include "mlb_ode.cuh"

namespace mlb {

#ifndef MLB_FN_ROLE_DIRS_WIDE
#define MLB_FN_ROLE_DIRS_WIDE 2
#endif
#ifndef MLB_FN_NG_WIDE
#define MLB_FN_NG_WIDE 6
#endif
template <int PARAM>
struct FnOde {
  static constexpr int U = 9, ND = 8, NPAIR = ND * (ND + 1) / 2, SIGMA = 6;
  static constexpr int NG = 6, PAIRS_PER_GROUP = NPAIR / NG;  // second-order pass split
  static constexpr int ROLE_DIRS = 2, N_ROLE = ND / ROLE_DIRS;  // Jacobian sweep split
  static constexpr int ROLE_DIRS_WIDE = MLB_FN_ROLE_DIRS_WIDE, WIDE_MIN_CHAINS = 8192;
  static constexpr int WIDE_MIN_CHAINS_SO = 8192;
  static constexpr int NG_WIDE = MLB_FN_NG_WIDE;
  static constexpr int GATE_WARPS = 1;  // no cooperative sweep: the RK4 stages are short
  template <int NS1, int NP>
  static constexpr int xch_doubles() { return 0; }
  static_assert(NPAIR % NG == 0, "pair groups must be equal-sized");

  struct Th {
    double a, b, g, d, e, z;  // alpha, beta, gamma, delta, epsilon, zeta
  };
  static MLB_DEV double loc(int k) {
    // fitzhugh_nagumo.py:20-28
    return k == 3 ? -1.0 : k == 4 ? -3.0 : k == 5 ? -2.0 : k == 6 ? -1.0 : 0.0;
  }
  static MLB_DEV double off(int k) { return PARAM == MLB_PARAM_NORMAL ? loc(k) : 0.0; }
  static MLB_DEV int dir_u(int d) { return d < 6 ? d : d + 1; }
  static MLB_DEV int u_dir(int a) { return a < 6 ? a : (a == 6 ? -1 : a - 1); }  // inverse
  static MLB_DEV Th theta(const double (&u)[U]) {
    Th t;
    t.a = exp(u[0] + off(0)), t.b = exp(u[1] + off(1)), t.g = exp(u[2] + off(2));
    t.d = exp(u[3] + off(3)), t.e = exp(u[4] + off(4)), t.z = exp(u[5] + off(5));
    return t;
  }
  static MLB_DEV double sigma(const double (&u)[U]) { return exp(u[6] + off(6)); }

  // prior nld of u (log-normal / normal priors pulled back to u): sum (u_k - m_k)^2 / 2
  static MLB_DEV double nld_u(const double (&u)[U], double (&g)[U]) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      double z = u[k] - ((PARAM == MLB_PARAM_NORMAL || k > 6) ? 0.0 : loc(k));
      g[k] = z;
      v = fma(0.5 * z, z, v);
    }
    return v;
  }

  // ---- right-hand side and its derivatives (fitzhugh_nagumo.py:35-42)
  static MLB_DEV void f(const Th& t, const double (&x)[2], double (&k)[2]) {
    k[0] = fma(t.a, x[0], fma(-t.b * x[0] * x[0], x[0], t.g * x[1]));
    k[1] = fma(-t.d, x[0], fma(-t.e, x[1], t.z));
  }
  // df/du_d at fixed x (theta_d = exp(u_d): d theta_d / d u_d = theta_d)
  template <int d>
  static MLB_DEV void phi(const Th& t, const double (&x)[2], double (&o)[2]) {
    o[0] = 0.0, o[1] = 0.0;
    if (d == 0) o[0] = t.a * x[0];
    if (d == 1) o[0] = -t.b * x[0] * x[0] * x[0];
    if (d == 2) o[0] = t.g * x[1];
    if (d == 3) o[1] = -t.d * x[0];
    if (d == 4) o[1] = -t.e * x[1];
    if (d == 5) o[1] = t.z;
  }
  // (d f_x / d u_d) s at fixed x
  template <int d>
  static MLB_DEV void dfx_du(const Th& t, const double (&x)[2], const double (&s)[2],
                             double (&o)[2]) {
    o[0] = 0.0, o[1] = 0.0;
    if (d == 0) o[0] = t.a * s[0];
    if (d == 1) o[0] = -3.0 * t.b * x[0] * x[0] * s[0];
    if (d == 2) o[0] = t.g * s[1];
    if (d == 3) o[1] = -t.d * s[0];
    if (d == 4) o[1] = -t.e * s[1];
  }

  // pair index k (0..35) <-> (a, b), a <= b, row-major over the upper triangle
  static __host__ __device__ constexpr int pair_a(int k) {
    int a = 0;
    while (k >= ND - a) k -= ND - a, ++a;
    return a;
  }
  static __host__ __device__ constexpr int pair_b(int k) {
    int a = 0;
    while (k >= ND - a) k -= ND - a, ++a;
    return a + k;
  }

  // One RK4 step of size h of the augmented system (x, S[ND], X[NP]) with low-storage
  // bookkeeping: y (state at the step start), acc (running RK4 combination) and ys (the
  // stage state, overwritten IN PLACE in dependency order second-order -> first-order ->
  // primal, so no separate copy of the stage derivative is kept).
  // S holds the NS1 first-order directions D0 .. D0+NS1-1 (a SUBSET when a Jacobian sweep
  // is split over several threads; second-order pairs need all of them: D0 = 0, NS1 = ND)
  template <int NS1, int NP, int P0>
  struct Aug {
    double x[2], S[NS1 ? NS1 : 1][2], X[NP ? NP : 1][2];
  };

  template <int NS1, int NP, int P0, bool first, bool last, int D0 = 0>
  static MLB_DEV void stage(const Th& t, const Aug<NS1, NP, P0>& y, Aug<NS1, NP, P0>& acc,
                            Aug<NS1, NP, P0>& ys, double w_acc, double c_next) {
    static_assert(NP == 0 || (D0 == 0 && NS1 == ND), "pairs need every direction");
    // Jacobian of f at the stage state
    const double a00 = fma(-3.0 * t.b * ys.x[0], ys.x[0], t.a), a01 = t.g, a10 = -t.d,
                 a11 = -t.e;
    if (NP > 0) {
      const double cxx = -6.0 * t.b * ys.x[0];
      [&]<int... K>(std::integer_sequence<int, K...>) {
        (
            [&] {
              constexpr int a = pair_a(P0 + K), b = pair_b(P0 + K);
              double r[2], m1[2], m2[2];
              dfx_du<b>(t, ys.x, ys.S[a], m1);
              dfx_du<a>(t, ys.x, ys.S[b], m2);
              r[0] = fma(cxx * ys.S[a][0], ys.S[b][0], m1[0] + m2[0]);
              r[1] = m1[1] + m2[1];
              if (a == b) {
                double p[2];
                phi<a>(t, ys.x, p);
                r[0] += p[0], r[1] += p[1];
              }
              const double k0 = fma(a00, ys.X[K][0], fma(a01, ys.X[K][1], r[0]));
              const double k1 = fma(a10, ys.X[K][0], fma(a11, ys.X[K][1], r[1]));
              acc.X[K][0] = fma(w_acc, k0, first ? y.X[K][0] : acc.X[K][0]);
              acc.X[K][1] = fma(w_acc, k1, first ? y.X[K][1] : acc.X[K][1]);
              if (!last) {
                ys.X[K][0] = fma(c_next, k0, y.X[K][0]);
                ys.X[K][1] = fma(c_next, k1, y.X[K][1]);
              }
            }(),
            ...);
      }(std::make_integer_sequence<int, NP>{});
    }
    if (NS1 > 0) {
      [&]<int... D>(std::integer_sequence<int, D...>) {
        (
            [&] {
              double p[2];
              phi<D0 + D>(t, ys.x, p);
              const double k0 = fma(a00, ys.S[D][0], fma(a01, ys.S[D][1], p[0]));
              const double k1 = fma(a10, ys.S[D][0], fma(a11, ys.S[D][1], p[1]));
              acc.S[D][0] = fma(w_acc, k0, first ? y.S[D][0] : acc.S[D][0]);
              acc.S[D][1] = fma(w_acc, k1, first ? y.S[D][1] : acc.S[D][1]);
              if (!last) {
                ys.S[D][0] = fma(c_next, k0, y.S[D][0]);
                ys.S[D][1] = fma(c_next, k1, y.S[D][1]);
              }
            }(),
            ...);
      }(std::make_integer_sequence<int, NS1>{});
    }
    double k[2];
    f(t, ys.x, k);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      acc.x[j] = fma(w_acc, k[j], first ? y.x[j] : acc.x[j]);
      if (!last) ys.x[j] = fma(c_next, k[j], y.x[j]);
    }
  }

  // Integrates over the schedule; obs(i, state) is called at every observation.
  // NS1 first-order directions starting at D0 (0 = primal only); NP pairs starting at P0.
  template <int NS1, int NP, int P0, class F, int D0 = 0, int GW = 1>
  static MLB_DEV void sweep(const OdeData& D, const double (&u)[U], F&& obs, double* = nullptr) {
    static_assert(GW == 1, "FitzHugh-Nagumo sweeps are not cooperative");
    const Th t = theta(u);
    Aug<NS1, NP, P0> y, acc, ys;
    y.x[0] = u[7], y.x[1] = u[8];
#pragma unroll
    for (int d = 0; d < (NS1 ? NS1 : 1); ++d) {
      y.S[d][0] = (D0 + d == 6) ? 1.0 : 0.0;
      y.S[d][1] = (D0 + d == 7) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < (NP ? NP : 1); ++k) y.X[k][0] = 0.0, y.X[k][1] = 0.0;
    int next = 0;
    int next_step = D.Y > 0 ? __ldg(D.obs_step) : -1;
#pragma unroll 1
    for (int s = 0; s < D.n_steps; ++s) {
      const double h = __ldg(D.dt_seq + s);
      ys = y;
      stage<NS1, NP, P0, true, false, D0>(t, y, acc, ys, h * (1.0 / 6.0), 0.5 * h);
      stage<NS1, NP, P0, false, false, D0>(t, y, acc, ys, h * (1.0 / 3.0), 0.5 * h);
      stage<NS1, NP, P0, false, false, D0>(t, y, acc, ys, h * (1.0 / 3.0), h);
      stage<NS1, NP, P0, false, true, D0>(t, y, acc, ys, h * (1.0 / 6.0), 0.0);
      y = acc;
      while (s == next_step) {
        obs(next, y);
        ++next;
        next_step = next < D.Y ? __ldg(D.obs_step + next) : -1;
      }
    }
  }
};

}  // namespace mlb
