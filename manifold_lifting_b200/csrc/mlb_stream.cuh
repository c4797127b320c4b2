// Streaming constrained-HMC device code for additive-noise models with LARGE dim_y
// (ODE-constrained models: FitzHugh-Nagumo Y = 200, U = 9): one chain per thread, the
// U-sized quantities (latent parameters, U x U capacitance matrix, its factor) in
// registers, every Y-sized quantity (Jacobian rows, residuals, noise coordinates,
// momenta) streamed through global memory in the component-major layout of the ABI, so a
// warp's access to row r of 32 consecutive chains is one coalesced 256 B transaction.
//
// Restates for this family, batched:
//   * IndependentAdditiveNoiseModelSystem, Woodbury branch (dim_u < dim_y):
//     mlift/systems.py:582-587 (J v, J^T w), 613-617 (capacitance Cholesky), 619-624
//     (Gram^-1 w), 626-633 (non-symmetric (J_l J_r^T)^-1 w, LU), 635-641 (log det)
//   * projection loops mlift/systems.py:190-228 / 230-270 / 272-326 and the wrappers'
//     classification mlift/solvers.py:81-95
//   * Mici's constrained leapfrog step
// Packed layouts (include/mlb.h): jac rows i*U+a (dy_du), then Y rows dy_dn;
// gram rows r*U+c (Cholesky factor of the capacitance matrix), then Y rows d = dy_dn^2.
#pragma once
#include "mlb_chmc.cuh"
#include "mlb_ode_fn.cuh"
#include "mlb_ode_hh.cuh"

namespace mlb {

// one chain's column of a [rows][ld] array
struct Col {
  double* p;
  int64_t ld;
  MLB_DEV double& operator[](int r) const { return p[(int64_t)r * ld]; }
  MLB_DEV Col rows(int r0) const { return Col{p + (int64_t)r0 * ld, ld}; }
  MLB_DEV explicit operator bool() const { return p != nullptr; }
};

// Per-chain scratch of the fused drivers, rows of one [rows][ld] allocation.
struct StreamScratch {
  Col Jc, Jt, Gc, c, w, g, qs, mu, cq, cp, dir;
  __host__ __device__ static int64_t rows(int U, int Y) {
    const int Q = U + Y, jac = U * Y + Y, gram = U * U + Y;
    return 2 * (int64_t)jac + gram + 3 * Y + 5 * Q;
  }
  MLB_DEV static StreamScratch make(double* base, int64_t ld, int64_t i, int U, int Y) {
    const int Q = U + Y, jac = U * Y + Y, gram = U * U + Y;
    StreamScratch s;
    Col b{base + i, ld};
    int r = 0;
    s.Jc = b.rows(r), r += jac;
    s.Jt = b.rows(r), r += jac;
    s.Gc = b.rows(r), r += gram;
    s.c = b.rows(r), r += Y;
    s.w = b.rows(r), r += Y;
    s.dir = b.rows(r), r += Y;
    s.g = b.rows(r), r += Q;
    s.qs = b.rows(r), r += Q;
    s.mu = b.rows(r), r += Q;
    s.cq = b.rows(r), r += Q;
    s.cp = b.rows(r), r += Q;
    return s;
  }
};

MLB_DEV double norm_accum(double acc, double v, int kind) {  // running norm of a stream
  if (kind == MLB_NORM_MAX) {
    unsigned long long a = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
    unsigned long long m = (unsigned long long)__double_as_longlong(acc);
    return __longlong_as_double((long long)(a > m ? a : m));
  }
  return fma(v, v, acc);
}
MLB_DEV double norm_finish(double acc, int kind) {
  return kind == MLB_NORM_MAX ? acc : sqrt(acc);
}

template <class M>
struct Stream {
  static constexpr int U = M::U;

  MLB_DEV static void load_u(const Col& q, double (&u)[U]) {
#pragma unroll
    for (int a = 0; a < U; ++a) u[a] = q[a];
  }

  // ---- model sweeps -----------------------------------------------------------------
  // c_i = x_i(u) + sigma(u) n_i - y_i with n_i = nfun(i); returns |c|
  template <class NF>
  MLB_DEV static double constr(const OdeData& D, const double (&u)[U], NF&& nfun, const Col& c,
                               int norm) {
    const double sig = M::sigma(u);
    double acc = 0.0;
    M::template sweep<0, 0, 0>(D, u, [&](int i, const auto& st) {
      const double ci = fma(sig, nfun(i), st.x[0] - __ldg(D.y_obs + i));
      if (c) c[i] = ci;
      acc = norm_accum(acc, ci, norm);
    });
    return norm_finish(acc, norm);
  }

  // Jacobian blocks + residual at q (systems.py:568-580); returns |c|
  MLB_DEV static double jacob(const OdeData& D, const Col& q, const Col& jac, const Col& c,
                              int norm) {
    double u[U];
    load_u(q, u);
    const double sig = M::sigma(u);
    const int Y = D.Y;
    double acc = 0.0;
    M::template sweep<M::ND, 0, 0>(D, u, [&](int i, const auto& st) {
      const double ni = q[U + i];
      const double ci = fma(sig, ni, st.x[0] - __ldg(D.y_obs + i));
      if (c) c[i] = ci;
      acc = norm_accum(acc, ci, norm);
#pragma unroll
      for (int d = 0; d < M::ND; ++d) jac[i * U + M::dir_u(d)] = st.S[d][0];
      jac[i * U + M::SIGMA] = sig * ni;  // d(sigma n_i)/du_sigma, sigma = exp(u + off)
      jac[U * Y + i] = sig;
    });
    return norm_finish(acc, norm);
  }

  // One ROLE of a Jacobian sweep split over threads: integrates the state plus the NDR
  // sensitivity directions D0 .. D0+NDR-1 and writes those dy_du columns; the role with
  // BASE also writes the residual, the sigma column and dy_dn (systems.py:568-580).
  // GW > 1: cooperative sweep of a (32, GW) CTA (see HhOde::integrate), xch its shared
  // exchange buffer; the callback (all stores) runs in warp 0 only.
  template <int D0, int NDR, bool BASE, int GW = 1>
  MLB_DEV static void jacob_role(const OdeData& D, const Col& q, const Col& jac, const Col& c,
                                 double* xch = nullptr) {
    double u[U];
    load_u(q, u);
    const double sig = M::sigma(u);
    const int Y = D.Y;
    auto cb = [&](int i, const auto& st) {
#pragma unroll
      for (int d = 0; d < NDR; ++d) jac[i * U + M::dir_u(D0 + d)] = st.S[d][0];
      if (BASE) {
        const double ni = q[U + i];
        if (c) c[i] = fma(sig, ni, st.x[0] - __ldg(D.y_obs + i));
        jac[i * U + M::SIGMA] = sig * ni;
        jac[U * Y + i] = sig;
      }
    };
    M::template sweep<NDR, 0, 0, decltype(cb)&, D0, GW>(D, u, cb, xch);
  }

  // ---- block algebra ----------------------------------------------------------------
  struct Chol {
    double L[U][U], rdiag[U];
    bool ok;
  };

  // capacitance matrix C = I + A^T D^-1 A, Cholesky (systems.py:613-617); writes packed gram
  MLB_DEV static bool decompose(int Y, const Col& jac, const Col& gram, Chol& C) {
#pragma unroll
    for (int a = 0; a < U; ++a)
#pragma unroll
      for (int b = 0; b < U; ++b) C.L[a][b] = (a == b) ? 1.0 : 0.0;
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      const double dn = jac[U * Y + i], d = dn * dn, rd = rcp_fast(d);
      if (gram) gram[U * U + i] = d;
      double row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[i * U + a];
#pragma unroll
      for (int a = 0; a < U; ++a) {
        const double ra = row[a] * rd;
#pragma unroll
        for (int b = 0; b <= a; ++b) C.L[a][b] = fma(ra, row[b], C.L[a][b]);
      }
    }
#pragma unroll
    for (int a = 0; a < U; ++a)
#pragma unroll
      for (int b = a + 1; b < U; ++b) C.L[a][b] = C.L[b][a];
    C.ok = chol_lower<U>(C.L, C.rdiag);
    if (gram) {
#pragma unroll
      for (int a = 0; a < U; ++a)
#pragma unroll
        for (int b = 0; b < U; ++b) gram[a * U + b] = C.L[a][b];
    }
    return C.ok;
  }

  MLB_DEV static void load_chol(const Col& gram, Chol& C) {
#pragma unroll
    for (int a = 0; a < U; ++a)
#pragma unroll
      for (int b = 0; b < U; ++b) C.L[a][b] = (b <= a) ? gram[a * U + b] : 0.0;
#pragma unroll
    for (int a = 0; a < U; ++a) C.rdiag[a] = rcp_fast(C.L[a][a]);
    C.ok = true;
  }

  // 0.5 log det(J J^T) = sum log diag chol(C) + 0.5 sum log d_i   (systems.py:635-641)
  MLB_DEV static double logdet(int Y, const Col& gram, const Chol& C) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < U; ++a) s += log(C.L[a][a]);
    double t = 0.0, prev = gram[U * U], lp = log(prev);
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {  // d is constant for scalar-noise models: reuse the log
      const double d = gram[U * U + i];
      if (d != prev) prev = d, lp = log(d);
      t += lp;
    }
    return fma(0.5, t, s);
  }

  // out = J v   (systems.py:582-584)
  MLB_DEV static void lmult(int Y, const Col& jac, const Col& v, const Col& out) {
    double vu[U];
    load_u(v, vu);
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      double s = jac[U * Y + i] * v[U + i];
#pragma unroll
      for (int a = 0; a < U; ++a) s = fma(jac[i * U + a], vu[a], s);
      out[i] = s;
    }
  }
  // out = J^T w   (systems.py:586-587)
  MLB_DEV static void rmult(int Y, const Col& jac, const Col& w, const Col& out) {
    double su[U];
#pragma unroll
    for (int a = 0; a < U; ++a) su[a] = 0.0;
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      const double wi = w[i];
#pragma unroll
      for (int a = 0; a < U; ++a) su[a] = fma(jac[i * U + a], wi, su[a]);
      out[U + i] = jac[U * Y + i] * wi;
    }
#pragma unroll
    for (int a = 0; a < U; ++a) out[a] = su[a];
  }

  // x = Gram^-1 w (systems.py:619-624), then out = J^T x if PINV else out = x.
  // w and the x result share storage (`wx`, Y rows).  If SUBTRACT, out -= J^T x instead.
  template <bool PINV, bool SUBTRACT>
  MLB_DEV static void inv_gram(int Y, const Col& jac, const Chol& C, const Col& wx,
                               const Col& out) {
    double t[U];
#pragma unroll
    for (int a = 0; a < U; ++a) t[a] = 0.0;
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      const double dn = jac[U * Y + i], wr = wx[i] * rcp_fast(dn * dn);
#pragma unroll
      for (int a = 0; a < U; ++a) t[a] = fma(jac[i * U + a], wr, t[a]);
    }
    chol_solve<U>(C.L, C.rdiag, t);
    double su[U];
#pragma unroll
    for (int a = 0; a < U; ++a) su[a] = 0.0;
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      const double dn = jac[U * Y + i];
      double s = wx[i];
      double row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[i * U + a], s = fma(-row[a], t[a], s);
      const double x = s * rcp_fast(dn * dn);
      if (PINV) {
#pragma unroll
        for (int a = 0; a < U; ++a) su[a] = fma(row[a], x, su[a]);
        if (SUBTRACT) out[U + i] -= dn * x; else out[U + i] = dn * x;
      } else {
        out[i] = x;
      }
    }
    if (PINV) {
#pragma unroll
      for (int a = 0; a < U; ++a) {
        if (SUBTRACT) out[a] -= su[a]; else out[a] = su[a];
      }
    }
  }

  // s = (I + A_r^T D_lr^-1 A_l)^-1 A_r^T D_lr^-1 w, D_lr = dn_l dn_r   (systems.py:626-632);
  // false if the LU factorisation met a zero pivot
  MLB_DEV static bool inv_jacprod_core(int Y, const Col& jl, const Col& jr, const Col& w,
                                       double (&s)[U]) {
    double mat[U][U];
#pragma unroll
    for (int a = 0; a < U; ++a) {
      s[a] = 0.0;
#pragma unroll
      for (int b = 0; b < U; ++b) mat[a][b] = (a == b) ? 1.0 : 0.0;
    }
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      const double rd = rcp_fast(jl[U * Y + i] * jr[U * Y + i]);
      const double wr = w[i] * rd;
      double rl[U];
#pragma unroll
      for (int b = 0; b < U; ++b) rl[b] = jl[i * U + b] * rd;
#pragma unroll
      for (int a = 0; a < U; ++a) {
        const double ra = jr[i * U + a];
        s[a] = fma(ra, wr, s[a]);
#pragma unroll
        for (int b = 0; b < U; ++b) mat[a][b] = fma(ra, rl[b], mat[a][b]);
      }
    }
    return lu_solve<U>(mat, s);
  }
  // x_i = (w_i - A_l[i] . s) / D_lr,i
  MLB_DEV static double inv_jacprod_x(int Y, int i, const Col& jl, const Col& jr, double wi,
                                      const double (&s)[U]) {
    double t = wi;
#pragma unroll
    for (int a = 0; a < U; ++a) t = fma(-jl[i * U + a], s[a], t);
    return t * rcp_fast(jl[U * Y + i] * jr[U * Y + i]);
  }

  // p -= J^T Gram^-1 J p   (systems.py:181-188, 464-466); w: Y rows of scratch
  MLB_DEV static void project(int Y, const Col& jac, const Chol& C, const Col& p, const Col& w) {
    lmult(Y, jac, p, w);
    inv_gram<true, true>(Y, jac, C, w, p);
  }

  // ---- evaluation at a position: everything Mici caches (systems.py:369-405) ---------
  struct EvalOut {
    double nld, logdet;
    bool ok;
  };
  // jac, gram <- blocks / factor at q; g <- dh1_dpos = grad nld + grad 0.5 log det Gram;
  // B: scratch with U*Y rows (holds B = Gram^-1 A during the second-order sweeps)
  MLB_DEV static EvalOut evaluate(const OdeData& D, const Col& q, const Col& jac,
                                  const Col& gram, const Col& g, const Col& B, Chol& C,
                                  const Col& c = Col{nullptr, 0}) {
    const int Y = D.Y;
    EvalOut e;
    jacob(D, q, jac, c, MLB_NORM_MAX);
    e.ok = decompose(Y, jac, gram, C);
    e.logdet = logdet(Y, gram, C);
    double u[U], gu[U];
    load_u(q, u);
    e.nld = M::nld_u(u, gu);
    grad_logdet(D, q, u, jac, C, B, gu, g, e.nld);
#pragma unroll
    for (int a = 0; a < U; ++a) g[a] = gu[a];
    return e;
  }

  // adds grad 0.5 log det(J J^T) to gu (latent part) and WRITES g rows U.. (noise part,
  // = n_i + B_i,sigma sigma) accumulating 0.5 |n|^2 into nld.  With M = C^-1 and
  // B = A M / sigma^2 (= Gram^-1 A):
  //   g_u[b]     += sum_i sum_a B_ia d2x_i/du_a du_b        (b an ODE direction)
  //   g_u[sigma] += Y - (U - tr M) + sum_i B_i,sigma sigma n_i
  //   g_n[i]      = n_i + B_i,sigma sigma
  template <bool WITH_SECOND_ORDER = true>
  MLB_DEV static void grad_logdet(const OdeData& D, const Col& q, const double (&u)[U],
                                  const Col& jac, const Chol& C, const Col& B,
                                  double (&gu)[U], const Col& g, double& nld) {
    const int Y = D.Y;
    double Minv[U][U];  // symmetric; computed column by column
#pragma unroll
    for (int j = 0; j < U; ++j) {
      double col[U];
#pragma unroll
      for (int a = 0; a < U; ++a) col[a] = (a == j) ? 1.0 : 0.0;
      chol_solve<U>(C.L, C.rdiag, col);
#pragma unroll
      for (int a = 0; a < U; ++a) Minv[a][j] = col[a];
    }
    double trM = 0.0;
#pragma unroll
    for (int a = 0; a < U; ++a) trM += Minv[a][a];
    double gs = (double)Y - ((double)U - trM), half_nn = 0.0;
#pragma unroll 4
    for (int i = 0; i < Y; ++i) {
      const double dn = jac[U * Y + i], rd = rcp_fast(dn * dn), ni = q[U + i];
      double row[U];
#pragma unroll
      for (int a = 0; a < U; ++a) row[a] = jac[i * U + a];
#pragma unroll
      for (int a = 0; a < U; ++a) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < U; ++k) s = fma(row[k], Minv[k][a], s);
        s *= rd;
        B[i * U + a] = s;
        if (a == M::SIGMA) {
          gs = fma(s, dn * ni, gs);
          g[U + i] = fma(s, dn, ni);
        }
      }
      half_nn = fma(0.5 * ni, ni, half_nn);
    }
    nld += half_nn;
    gu[M::SIGMA] += gs;
    if (!WITH_SECOND_ORDER) return;
    // second-order sensitivity sweeps, NG groups of pairs (a <= b)
    [&]<int... G>(std::integer_sequence<int, G...>) {
      (second_order_group<G>(D, u, B, gu), ...);
    }(std::make_integer_sequence<int, M::NG>{});
  }

  // NP pairs per group (default: the model's NG groups; the wide form of the phased driver
  // takes more pairs per thread, see kp_second_order)
  template <int G, int GW = 1, int NP = M::PAIRS_PER_GROUP>
  MLB_DEV static void second_order_group(const OdeData& D, const double (&u)[U], const Col& B,
                                         double (&gu)[U], double* xch = nullptr) {
    constexpr int P0 = G * NP;
    double h[M::ND];
#pragma unroll
    for (int d = 0; d < M::ND; ++d) h[d] = 0.0;
    auto cb = [&](int i, const auto& st) {
      double b[M::ND];
#pragma unroll
      for (int d = 0; d < M::ND; ++d) b[d] = B[i * U + M::dir_u(d)];
      [&]<int... K>(std::integer_sequence<int, K...>) {
        (
            [&] {
              constexpr int a = M::pair_a(P0 + K), c = M::pair_b(P0 + K);
              h[c] = fma(b[a], st.X[K][0], h[c]);
              if (a != c) h[a] = fma(b[c], st.X[K][0], h[a]);
            }(),
            ...);
      }(std::make_integer_sequence<int, NP>{});
    };
    M::template sweep<M::ND, NP, P0, decltype(cb)&, 0, GW>(D, u, cb, xch);
#pragma unroll
    for (int d = 0; d < M::ND; ++d) gu[M::dir_u(d)] += h[d];
  }

  // ---- projection solvers (systems.py:190-326), q in/out, mu_dt out (Q rows) ----------
  template <int SOLVER>
  MLB_DEV static SolveOut solve(const OdeData& D, const Col& q, const Col& jp, const Chol& Cp,
                                double dt, const Opts& o, const Col& mu, const Col& jt,
                                const Col& c, const Col& dir) {
    const int Y = D.Y, Q = U + Y;
    SolveOut r;
    r.iters = 0, r.n_constr = 0, r.status = MLB_ST_OK, r.ndq = INFINITY, r.err = -1.0;
#pragma unroll 1
    for (int k = 0; k < Q; ++k) mu[k] = 0.0;
#pragma unroll 1
    while (keep_iterating(r.iters, r.ndq, r.err, o)) {
      double u[U], s[U], du[U];
      load_u(q, u);
#pragma unroll
      for (int a = 0; a < U; ++a) du[a] = 0.0;
      double nacc = 0.0;
      if (SOLVER == MLB_SOLVER_QUASI_NEWTON) {
        r.err = constr(D, u, [&](int i) { return q[U + i]; }, c, o.norm);
        // delta = J_prev^T Gram_prev^-1 c
#pragma unroll
        for (int a = 0; a < U; ++a) s[a] = 0.0;
#pragma unroll 4
        for (int i = 0; i < Y; ++i) {
          const double dn = jp[U * Y + i], cr = c[i] * rcp_fast(dn * dn);
#pragma unroll
          for (int a = 0; a < U; ++a) s[a] = fma(jp[i * U + a], cr, s[a]);
        }
        chol_solve<U>(Cp.L, Cp.rdiag, s);
#pragma unroll 4
        for (int i = 0; i < Y; ++i) {
          const double x = inv_jacprod_x(Y, i, jp, jp, c[i], s);
#pragma unroll
          for (int a = 0; a < U; ++a) du[a] = fma(jp[i * U + a], x, du[a]);
          const double dni = jp[U * Y + i] * x;
          q[U + i] -= dni, mu[U + i] += dni;
          nacc = norm_accum(nacc, dni, o.norm);
        }
      } else {
        const double err_here = jacob(D, q, jt, c, o.norm);
        if (!inv_jacprod_core(Y, jt, jp, c, s)) {
#pragma unroll
          for (int a = 0; a < U; ++a) s[a] = nan("");
        }
        if (SOLVER == MLB_SOLVER_NEWTON) {
          r.err = err_here;
#pragma unroll 4
          for (int i = 0; i < Y; ++i) {
            const double x = inv_jacprod_x(Y, i, jt, jp, c[i], s);
#pragma unroll
            for (int a = 0; a < U; ++a) du[a] = fma(jp[i * U + a], x, du[a]);
            const double dni = jp[U * Y + i] * x;
            q[U + i] -= dni, mu[U + i] += dni;
            nacc = norm_accum(nacc, dni, o.norm);
          }
        } else {  // Newton with backtracking line search (systems.py:272-326)
#pragma unroll 4
          for (int i = 0; i < Y; ++i) {
            const double x = inv_jacprod_x(Y, i, jt, jp, c[i], s);
#pragma unroll
            for (int a = 0; a < U; ++a) du[a] = fma(jp[i * U + a], x, du[a]);
            dir[i] = jp[U * Y + i] * x;
          }
          double step = 1.0, used = 1.0, new_err;
          int j = 0;
          do {  // first trial unconditional, then halve while the residual grew
            double ut[U];
#pragma unroll
            for (int a = 0; a < U; ++a) ut[a] = fma(-step, du[a], u[a]);
            new_err = constr(D, ut, [&](int i) { return fma(-step, dir[i], q[U + i]); },
                             Col{nullptr, 0}, o.norm);
            used = step;
            step *= 0.5;
            j += 1;
          } while (j < o.max_ls && new_err > err_here);
#pragma unroll 4
          for (int i = 0; i < Y; ++i) {
            const double dni = used * dir[i];
            q[U + i] -= dni, mu[U + i] += dni;
            nacc = norm_accum(nacc, dni, o.norm);
          }
#pragma unroll
          for (int a = 0; a < U; ++a) du[a] *= used;
          r.err = new_err;
          r.n_constr += j;
        }
      }
#pragma unroll
      for (int a = 0; a < U; ++a) {
        q[a] = u[a] - du[a];
        mu[a] += du[a];
        nacc = norm_accum(nacc, du[a], o.norm);
      }
      r.ndq = norm_finish(nacc, o.norm);
      r.iters += 1;
    }
    const double rdt = 1.0 / dt;
#pragma unroll 1
    for (int k = 0; k < Q; ++k) mu[k] *= rdt;
    r.status = classify(r.err, r.ndq, o);
    return r;
  }

  // ---- fused trajectory (same sub-phase loop as run_trajectory in mlb_chmc.cuh, with
  // the chain state in global memory) ------------------------------------------------
  template <int SOLVER, bool PROJECT_FIRST, bool CHECK_H>
  MLB_DEV static TrajOut trajectory(const OdeData& D, const Col& qw, const Col& pw,
                                    const StreamScratch& S, int n_step, double dt,
                                    const Opts& o, double max_delta_h) {
    const int Y = D.Y, Q = U + Y;
    TrajOut t;
    t.status = MLB_ST_OK, t.done = 0, t.it_fwd = 0, t.it_rev = 0;
    t.h_init = t.h_last = nan("");
    Chol C;
    EvalOut e;
    e.nld = e.logdet = 0.0, e.ok = true;
    int sub = PROJECT_FIRST ? -1 : 0;
    bool first = true;
#pragma unroll 1
    for (int k = 0; k < Q; ++k) S.cq[k] = qw[k], S.cp[k] = pw[k];
    auto hamiltonian = [&]() {
      double s = 0.0;
#pragma unroll 1
      for (int k = 0; k < Q; ++k) s = fma(pw[k], pw[k], s);
      return e.nld + e.logdet + 0.5 * s;
    };
#pragma unroll 1
    for (;;) {
      if (first || sub == 1) {
        e = evaluate(D, qw, S.Jc, S.Gc, S.g, S.Jt, C);
        if (first && !PROJECT_FIRST) t.h_init = t.h_last = hamiltonian();
        first = false;
        if (!e.ok) {
          t.status = MLB_ST_LINALG;
          break;
        }
        if (!PROJECT_FIRST && n_step <= 0) break;
      }
      if (sub == 0 || sub == 2) {
#pragma unroll 1
        for (int k = 0; k < Q; ++k) pw[k] = fma(-0.5 * dt, S.g[k], pw[k]);
      }
      project(Y, S.Jc, C, pw, S.w);
      if (PROJECT_FIRST && sub == -1) {
        t.h_init = t.h_last = hamiltonian();
#pragma unroll 1
        for (int k = 0; k < Q; ++k) S.cp[k] = pw[k];
        if (n_step <= 0) break;
        sub = 0;
        continue;
      }
      if (sub == 2) {
#pragma unroll 1
        for (int k = 0; k < Q; ++k) S.cq[k] = qw[k], S.cp[k] = pw[k];
        t.h_last = hamiltonian();
        t.done += 1;
        if (CHECK_H && fabs(t.h_last - t.h_init) > max_delta_h) {
          t.status = MLB_ST_H_DIVERGED;
          break;
        }
        if (t.done >= n_step) break;
        sub = 0;
        continue;
      }
      const double sdt = (sub == 0) ? dt : -dt;
#pragma unroll 1
      for (int k = 0; k < Q; ++k) S.qs[k] = fma(sdt, pw[k], qw[k]);
      const SolveOut r = solve<SOLVER>(D, S.qs, S.Jc, C, sdt, o, S.mu, S.Jt, S.c, S.dir);
      if (sub == 0) t.it_fwd += r.iters; else t.it_rev += r.iters;
      if (r.status != MLB_ST_OK) {
        t.status = r.status;
        break;
      }
      if (sub == 0) {
#pragma unroll 1
        for (int k = 0; k < Q; ++k) pw[k] -= S.mu[k], qw[k] = S.qs[k];
      } else {
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) acc = norm_accum(acc, S.qs[k] - S.cq[k], o.rev_norm);
        if (norm_finish(acc, o.rev_norm) > o.rev_tol) {
          t.status = MLB_ST_NON_REVERSIBLE;
          break;
        }
      }
      sub += 1;
    }
    if (t.status != MLB_ST_OK && t.status != MLB_ST_H_DIVERGED) {
#pragma unroll 1
      for (int k = 0; k < Q; ++k) qw[k] = S.cq[k], pw[k] = S.cp[k];
    }
    return t;
  }
};

}  // namespace mlb
