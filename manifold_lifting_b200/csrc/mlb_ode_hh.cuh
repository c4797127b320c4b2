// Hodgkin-Huxley current-stimulus model: exponential-Euler sweeps with forward-mode jets,
// one chain per thread.
//
// Reference: mlift/example_models/hh_current_stimulus.py:19-27 (priors), 34-114 (rate
// functions), 125-139 (steady-state initial condition at v = E_leak), 142-178
// (`integrate_ode`: v, n, m, h, p advanced by exponential-Euler steps over
// linspace(0, end_time, n)), 181-197 (observation of v inside the stimulus window plus
// sigma n).  u = (k_tau_p_1, g_bar_Na, g_bar_K, g_bar_M, g_leak, v_t, sigma).  The time
// grid, the per-step stimulus current (:117-122) and the observation steps are computed on
// the host from the reference's expressions and are model data.
//
// The reference differentiates the traced recurrence with jax.jacfwd and
// reverse-over-forward autodiff (systems.py:332-334, 568-580).  Here the recurrence is
// written once for a scalar type T and run with T = double (residual), first-order jets in
// the six ODE directions (Jacobian) and jets carrying three second derivatives at a time
// (gradient of the log-det term; 21 pairs in 7 groups, first-order part recomputed per
// group so that a group fits the register file).
#pragma once
#include "mlb_jet.cuh"
#include "mlb_ode.cuh"

namespace mlb {

// indices into OdeData::consts (first 27 in the order of the reference's data file
// constants, then reciprocals precomputed on the host)
enum {
  HH_k_alpha_n_1, HH_k_alpha_n_2, HH_k_alpha_n_3, HH_k_beta_n_1, HH_k_beta_n_2, HH_k_beta_n_3,
  HH_k_alpha_m_1, HH_k_alpha_m_2, HH_k_alpha_m_3, HH_k_beta_m_1, HH_k_beta_m_2, HH_k_beta_m_3,
  HH_k_alpha_h_1, HH_k_alpha_h_2, HH_k_alpha_h_3, HH_k_beta_h_1, HH_k_beta_h_2, HH_k_beta_h_3,
  HH_k_p_infinity_1, HH_k_p_infinity_2, HH_k_tau_p_2, HH_k_tau_p_3, HH_k_tau_p_4, HH_E_Na,
  HH_E_K, HH_E_leak, HH_c_m,
  HH_r_alpha_n_3, HH_r_beta_n_3, HH_r_alpha_m_3, HH_r_beta_m_3, HH_r_alpha_h_3, HH_r_beta_h_3,
  HH_r_p_infinity_2, HH_r_tau_p_4, HH_r_c_m, HH_NCONST
};
static_assert(HH_NCONST <= kOdeConsts, "constants array too small");

#ifndef MLB_HH_ROLE_DIRS
#define MLB_HH_ROLE_DIRS 1
#endif
#ifndef MLB_HH_ROLE_DIRS_WIDE
#define MLB_HH_ROLE_DIRS_WIDE 3
#endif
#ifndef MLB_HH_NG
#define MLB_HH_NG 7
#endif
#ifndef MLB_HH_GATE_WARPS
#define MLB_HH_GATE_WARPS 4
#endif
template <int PARAM>
struct HhOde {
  static constexpr int U = 7, ND = 6, NPAIR = ND * (ND + 1) / 2, SIGMA = 6;
  static constexpr int NG = MLB_HH_NG, PAIRS_PER_GROUP = NPAIR / NG;
  static constexpr int ROLE_DIRS = MLB_HH_ROLE_DIRS, N_ROLE = ND / ROLE_DIRS;  // Jacobian sweep split
  // launches of more than WIDE_MIN_CHAINS chains take ROLE_DIRS_WIDE directions per role
  // (a threshold of 1,536 - the solver rounds carry chains + twins since the reversibility
  // checks run one step behind - was tried: 107 instead of 143 ms per 4 steps for ONE
  // 1,024-chain share of the 8-GPU split, but 134 / 125 instead of 116 / 120 ms for two
  // others and 176 instead of 146 ms on eight GPUs (max over ranks): not kept,
  // profiles/r7b_hh_wide.txt, r7e_hh_ranks.txt); the second-order sweep has its own threshold
  static constexpr int ROLE_DIRS_WIDE = MLB_HH_ROLE_DIRS_WIDE, WIDE_MIN_CHAINS = 3072;
  static constexpr int WIDE_MIN_CHAINS_SO = 3072;
  // second-order sweep of a full GPU: all 21 pairs per thread.  At 8,192 chains that is 256
  // warps on 592 schedulers (11.1 ms, 1.7 warps active per SM,
  // profiles/r6j_full_kp_second_order_hh.md); three groups of 7 pairs (768 warps, each
  // repeating the state + first-order part) measured the same per step: 302.8 against
  // 297.9 ms per 4 steps (MLB_HH_NG_WIDE=3)
#ifndef MLB_HH_NG_WIDE
#define MLB_HH_NG_WIDE 1
#endif
  static constexpr int NG_WIDE = MLB_HH_NG_WIDE;
  static constexpr int GATE_WARPS = MLB_HH_GATE_WARPS;  // warps sharing one chain batch in a cooperative sweep
  static_assert(NPAIR % NG == 0, "pair groups must be equal-sized");
  static MLB_DEV int dir_u(int d) { return d; }
  static MLB_DEV int u_dir(int a) { return a < ND ? a : -1; }  // inverse (-1: sigma)
  static __host__ __device__ constexpr int pair_a(int k) { return PairMap<ND>::a(k); }
  static __host__ __device__ constexpr int pair_b(int k) { return PairMap<ND>::b(k); }

  // hh_current_stimulus.py:19-27
  static MLB_DEV double loc(int k) {
    return k == 0 ? 8.0 : k == 1 ? 4.0 : k == 2 ? 2.0 : k == 3 ? -3.0 : k == 4 ? -3.0
         : k == 5 ? -60.0 : 0.0;
  }
  static MLB_DEV double scale(int k) { return k == 5 ? 10.0 : 1.0; }
  static MLB_DEV double sigma(const double (&u)[U]) {
    return exp(u[6] + (PARAM == MLB_PARAM_NORMAL ? loc(6) : 0.0));
  }
  static MLB_DEV double nld_u(const double (&u)[U], double (&g)[U]) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const double rs = (PARAM == MLB_PARAM_NORMAL) ? 1.0 : 1.0 / scale(k);
      const double z = (PARAM == MLB_PARAM_NORMAL) ? u[k] : (u[k] - loc(k)) * rs;
      g[k] = z * rs;
      v = fma(0.5 * z, z, v);
    }
    return v;
  }

  // ---- scalar-type plumbing: T is double or DJet<ND, NP, P0>
  template <class T>
  using Ops = JetOps<T>;

  template <class T>
  struct Par {
    T k_tau_p_1, g_Na, g_K, g_M, g_leak, v_t;
  };

  // D0: first direction carried by T's gradient (a jet may carry a SUBSET of the six ODE
  // directions when a Jacobian sweep is split over threads; seeds outside it vanish)
  template <class T, int D0 = 0>
  static MLB_DEV Par<T> params(const double (&u)[U]) {
    T t[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (k < 5) {
        t[k] = jexp(Ops<T>::seed(u[k] + (PARAM == MLB_PARAM_NORMAL ? loc(k) : 0.0), k - D0, 1.0));
      } else {
        t[k] = (PARAM == MLB_PARAM_NORMAL)
                   ? Ops<T>::seed(fma(scale(k), u[k], loc(k)), k - D0, scale(k))
                   : Ops<T>::seed(u[k], k - D0, 1.0);
      }
    }
    return Par<T>{t[0], t[1], t[2], t[3], t[4], t[5]};
  }

  template <class T>
  static MLB_DEV T xoe(const T& x) { return x * jrecip(jexpm1(x)); }  // x / expm1(x), :34-35

  template <class T>
  struct Rates {
    T an, bn, am, bm, ah, bh;
  };
  // opening / closing rates of one gate (GATE 0: n, 1: m, 2: h) at potential v
  template <int GATE, class T>
  static MLB_DEV void gate_rates(const OdeData& D, const T& v, const Par<T>& p, T& a, T& b) {
    const double* K = D.consts;
    const T w = v - p.v_t;
    if constexpr (GATE == 0) {
      a = (K[HH_k_alpha_n_1] * K[HH_k_alpha_n_3]) * xoe((w - K[HH_k_alpha_n_2]) * (-K[HH_r_alpha_n_3]));
      b = K[HH_k_beta_n_1] * jexp((w - K[HH_k_beta_n_2]) * (-K[HH_r_beta_n_3]));
    } else if constexpr (GATE == 1) {
      a = (K[HH_k_alpha_m_1] * K[HH_k_alpha_m_3]) * xoe((w - K[HH_k_alpha_m_2]) * (-K[HH_r_alpha_m_3]));
      b = (K[HH_k_beta_m_1] * K[HH_k_beta_m_3]) * xoe((w - K[HH_k_beta_m_2]) * K[HH_r_beta_m_3]);
    } else {
      a = K[HH_k_alpha_h_1] * jexp((w - K[HH_k_alpha_h_2]) * (-K[HH_r_alpha_h_3]));
      b = K[HH_k_beta_h_1] * jrecip(jexp((w - K[HH_k_beta_h_2]) * (-K[HH_r_beta_h_3])) + 1.0);
    }
  }
  template <class T>
  static MLB_DEV Rates<T> rates(const OdeData& D, const T& v, const Par<T>& p) {
    Rates<T> r;
    gate_rates<0>(D, v, p, r.an, r.bn);
    gate_rates<1>(D, v, p, r.am, r.bm);
    gate_rates<2>(D, v, p, r.ah, r.bh);
    return r;
  }
  template <class T>
  static MLB_DEV T p_infinity(const OdeData& D, const T& v) {
    const double* K = D.consts;
    return jrecip(1.0 + jexp((K[HH_k_p_infinity_1] - v) * K[HH_r_p_infinity_2]));
  }

  // GW = 1: one thread advances the whole recurrence.  GW = GATE_WARPS (= 4): the CTA is
  // (32 chains) x (4 warps); every warp advances the membrane potential (it needs all four
  // gating variables) but only ITS OWN gate (threadIdx.y: n, m, h, p), whose new value it
  // publishes through shared memory `xch` ([2][4][JetSize][32] doubles, double-buffered:
  // one __syncthreads per step).  A step's critical path drops from 10 exp + 3 expm1 + 11
  // reciprocals to those of the potential plus one gate (sweeps of a 1,024-chain batch are
  // latency-bound: 32 warps x roles on 592 schedulers).  obs is called by warp 0 only.
  template <class T, class F, int D0 = 0, int GW = 1>
  static MLB_DEV void integrate(const OdeData& D, const double (&u)[U], F&& obs,
                                double* xch = nullptr) {
    const double* K = D.consts;
    const Par<T> p = params<T, D0>(u);
    const int gate = GW > 1 ? (int)threadIdx.y : 0;
    constexpr int JS = JetSize<T>::value;
    T v = Ops<T>::constant(K[HH_E_leak]);
    T n, m, h, pp;
    {  // steady state at v = E_leak (:125-139)
      const Rates<T> r = rates<T>(D, v, p);
      n = r.an * jrecip(r.an + r.bn);
      m = r.am * jrecip(r.am + r.bm);
      h = r.ah * jrecip(r.ah + r.bh);
      pp = p_infinity<T>(D, v);
    }
    int next = 0, buf = 0;
    int next_step = D.Y > 0 ? __ldg(D.obs_step) : -1;
#pragma unroll 1
    for (int s = 0; s < D.n_steps && next < D.Y; ++s) {
      const double dt = __ldg(D.dt_seq + s), stim = __ldg(D.aux + s);
      // :145-160  membrane potential
      const T n2 = n * n;
      const T g_K = p.g_K * (n2 * n2);
      const T g_Na = p.g_Na * ((m * m) * m) * h;
      const T g_M = p.g_M * pp;
      const T g_tot = g_K + g_Na + g_M + p.g_leak;
      const T tau_v = K[HH_c_m] * jrecip(g_tot);
      const T v_inf = ((g_K + g_M) * K[HH_E_K] + g_Na * K[HH_E_Na] + p.g_leak * K[HH_E_leak] + stim) *
                      tau_v * K[HH_r_c_m];
      const T v_ = v_inf + (v - v_inf) * jexp(g_tot * (-dt * K[HH_r_c_m]));  // exp(-dt / tau_v)
      // :161-173  gates, rates evaluated at the NEW potential
      if (GW == 1 || gate == 0) {
        T a, b;
        gate_rates<0>(D, v_, p, a, b);
        const T sn = a + b, n_inf = a * jrecip(sn);
        n = n_inf + (n - n_inf) * jexp(sn * (-dt));  // tau_n = 1 / (a_n + b_n)
      }
      if (GW == 1 || gate == 1) {
        T a, b;
        gate_rates<1>(D, v_, p, a, b);
        const T sm = a + b, m_inf = a * jrecip(sm);
        m = m_inf + (m - m_inf) * jexp(sm * (-dt));
      }
      if (GW == 1 || gate == 2) {
        T a, b;
        gate_rates<2>(D, v_, p, a, b);
        const T sh = a + b, h_inf = a * jrecip(sh);
        h = h_inf + (h - h_inf) * jexp(sh * (-dt));
      }
      if (GW == 1 || gate == 3) {
        const T p_inf = p_infinity<T>(D, v_);
        // tau_p = k_tau_p_1 / (k2 e + 1 / e), e = exp((v + k3) / k4)   (:104-114)
        const T e = jexp((v_ + K[HH_k_tau_p_3]) * K[HH_r_tau_p_4]);
        const T rate = (K[HH_k_tau_p_2] * e + jrecip(e)) * jrecip(p.k_tau_p_1);
        pp = p_inf + (pp - p_inf) * jexp(rate * (-dt));
      }
      if constexpr (GW > 1) {
        double* mine = xch + ((buf * GW + gate) * JS) * 32 + threadIdx.x;
        if (gate == 0) jet_store(n, mine, 32);
        else if (gate == 1) jet_store(m, mine, 32);
        else if (gate == 2) jet_store(h, mine, 32);
        else jet_store(pp, mine, 32);
        __syncthreads();
        const double* all = xch + (buf * GW * JS) * 32 + threadIdx.x;
        if (gate != 0) jet_load(n, all + 0 * JS * 32, 32);
        if (gate != 1) jet_load(m, all + 1 * JS * 32, 32);
        if (gate != 2) jet_load(h, all + 2 * JS * 32, 32);
        if (gate != 3) jet_load(pp, all + 3 * JS * 32, 32);
        buf ^= 1;
      }
      v = v_;
      while (s == next_step) {
        if (GW == 1 || gate == 0) obs(next, v);
        ++next;
        next_step = next < D.Y ? __ldg(D.obs_step + next) : -1;
      }
    }
  }

  template <int NS1, int NP>
  struct ObsView {
    double x[1], S[NS1 ? NS1 : 1][1], X[NP ? NP : 1][1];
  };

  // same interface as FnOde::sweep (see mlb_stream.cuh): NS1 first-order directions
  // starting at D0 (0 = value only), NP second-order pairs starting at pair index P0
  // GW > 1: cooperative form for a (32, GW) CTA, xch = shared exchange buffer of
  // xch_doubles<NS1, NP>() doubles (see integrate); obs runs in warp 0 only
  template <int NS1, int NP>
  static constexpr int xch_doubles() {
    return 2 * GATE_WARPS * (1 + NS1 + NP) * 32;
  }
  template <int NS1, int NP, int P0, class F, int D0 = 0, int GW = 1>
  static MLB_DEV void sweep(const OdeData& D, const double (&u)[U], F&& obs,
                            double* xch = nullptr) {
    static_assert(NP == 0 || (D0 == 0 && NS1 == ND), "pairs need every direction");
    if constexpr (NS1 == 0) {
      auto cb = [&](int i, const double& v) {
        ObsView<0, 0> o;
        o.x[0] = v;
        obs(i, o);
      };
      integrate<double, decltype(cb)&, 0, GW>(D, u, cb, xch);
    } else {
      using T = DJet<NS1, NP, P0>;
      auto cb = [&](int i, const T& v) {
        ObsView<NS1, NP> o;
        o.x[0] = v.v;
#pragma unroll
        for (int d = 0; d < NS1; ++d) o.S[d][0] = v.g[d];
#pragma unroll
        for (int k = 0; k < NP; ++k) o.X[k][0] = v.h[k];
        obs(i, o);
      };
      integrate<T, decltype(cb)&, D0, GW>(D, u, cb, xch);
    }
  }
};

}  // namespace mlb
