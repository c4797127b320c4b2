// CPU oracle (closed-form, scalar, one chain at a time) for the constrained-HMC hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
// baseline legs may load the library built from this file; the product never does.
// PARITY UNPINNED: the reference has no fixtures for this path and cannot be run here
// (no JAX / Mici); this restatement is cross-checked against the independent torch
// autodiff restatement in oracle/mlift_oracle and against the invariants of SURVEY 8c.
//
// What is restated (paths under /root/reference/):
//   * generic block-structured constraint algebra, dense-Cholesky and Woodbury
//     branches                       mlift/systems.py:582-641 (additive noise)
//                                    mlift/systems.py:712-794 (hierarchical)
//   * default pinv / normal-space compositions   mlift/systems.py:173-188
//   * the three projection loops     mlift/systems.py:190-228, 230-270, 272-326
//   * solver success / failure classification    mlift/solvers.py:81-95
//   * norms                          mlift/solvers.py:6-13
//   * h1, dh1_dpos, h2, h2_flow, cotangent projection  mlift/systems.py:407-466
//   * eight-schools model            mlift/example_models/eight_schools.py:28-53
//     with priors/transforms         mlift/prior.py:17-55, distributions.py:141-168,
//                                    439-465, transforms.py:43-55,87-95,137-144
//   * toy model                      mlift/example_models/toy_2d_model.py:43-87 (+ notebook)
//   * Mici 0.2.1 constrained leapfrog step, Mici Newton / quasi-Newton solvers, static
//     Metropolis accept rule: NOT in the reference tree, restated from published
//     behaviour (notebook cell `simulate_proposal`).
//
// Layout: chains are rows, AoS [n_chain][dim] row-major (NumPy natural).  OpenMP over
// chains is used only by the batch drivers (the CPU baseline), never inside a chain.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_models.hpp"

namespace orc {

// ----------------------------------------------------------------- dense helpers

// in-place lower Cholesky of row-major K x K; returns false if not positive definite
static bool chol_lower(double* a, int K) {
  for (int j = 0; j < K; ++j) {
    double s = a[j * K + j];
    for (int k = 0; k < j; ++k) s -= a[j * K + k] * a[j * K + k];
    if (!(s > 0.0)) return false;
    double d = std::sqrt(s);
    a[j * K + j] = d;
    for (int i = j + 1; i < K; ++i) {
      double t = a[i * K + j];
      for (int k = 0; k < j; ++k) t -= a[i * K + k] * a[j * K + k];
      a[i * K + j] = t / d;
    }
    for (int i = 0; i < j; ++i) a[i * K + j] = 0.0;
  }
  return true;
}

static void chol_solve(const double* L, int K, double* b) {
  for (int i = 0; i < K; ++i) {
    double t = b[i];
    for (int k = 0; k < i; ++k) t -= L[i * K + k] * b[k];
    b[i] = t / L[i * K + i];
  }
  for (int i = K - 1; i >= 0; --i) {
    double t = b[i];
    for (int k = i + 1; k < K; ++k) t -= L[k * K + i] * b[k];
    b[i] = t / L[i * K + i];
  }
}

// LU with partial pivoting (what np.linalg.solve / LAPACK gesv does); a destroyed
static bool lu_solve(double* a, int K, double* b) {
  for (int j = 0; j < K; ++j) {
    int piv = j;
    double best = std::fabs(a[j * K + j]);
    for (int i = j + 1; i < K; ++i)
      if (std::fabs(a[i * K + j]) > best) best = std::fabs(a[i * K + j]), piv = i;
    if (!(best > 0.0)) return false;
    if (piv != j) {
      for (int k = 0; k < K; ++k) std::swap(a[j * K + k], a[piv * K + k]);
      std::swap(b[j], b[piv]);
    }
    for (int i = j + 1; i < K; ++i) {
      double f = a[i * K + j] / a[j * K + j];
      a[i * K + j] = f;
      for (int k = j + 1; k < K; ++k) a[i * K + k] -= f * a[j * K + k];
      b[i] -= f * b[j];
    }
  }
  for (int i = K - 1; i >= 0; --i) {
    double t = b[i];
    for (int k = i + 1; k < K; ++k) t -= a[i * K + k] * b[k];
    b[i] = t / a[i * K + i];
  }
  return true;
}

static double vec_norm(const double* v, int n, int kind) {
  if (kind == 0) {  // maximum norm, NaN-propagating like numpy's max
    double m = 0.0;
    for (int i = 0; i < n; ++i) {
      double a = std::fabs(v[i]);
      if (std::isnan(a)) return a;
      if (a > m) m = a;
    }
    return m;
  }
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return std::sqrt(s);
}

// ----------------------------------------------------------- generic block algebra

struct Jac {  // constraint Jacobian blocks at one point
  std::vector<double> dy_du, dy_dv, dy_dn;
  void init(const Model& M) {
    dy_du.assign((size_t)M.Y * M.U, 0.0);
    dy_dv.assign(M.family == FAMILY_HIER ? M.Y : 0, 0.0);
    dy_dn.assign(M.Y, 0.0);
  }
};

struct Gram {  // dense: chol [Y x Y]; Woodbury: chol of capacitance [U x U] and d [Y]
  std::vector<double> chol, d;
  void init(const Model& M) {
    int K = M.dense ? M.Y : M.U;
    chol.assign((size_t)K * K, 0.0);
    d.assign(M.Y, 0.0);
  }
};

static void lmult_jacob(const Model& M, const Jac& J, const double* v, double* out) {
  const int U = M.U, Y = M.Y;
  for (int i = 0; i < Y; ++i) {
    double s = 0.0;
    for (int a = 0; a < U; ++a) s += J.dy_du[(size_t)i * U + a] * v[a];
    if (M.family == FAMILY_HIER)
      s += J.dy_dv[i] * v[U + i] + J.dy_dn[i] * v[U + Y + i];
    else
      s += J.dy_dn[i] * v[U + i];
    out[i] = s;
  }
}

static void rmult_jacob(const Model& M, const Jac& J, const double* w, double* out) {
  const int U = M.U, Y = M.Y;
  for (int a = 0; a < U; ++a) {
    double s = 0.0;
    for (int i = 0; i < Y; ++i) s += w[i] * J.dy_du[(size_t)i * U + a];
    out[a] = s;
  }
  if (M.family == FAMILY_HIER) {
    for (int i = 0; i < Y; ++i) out[U + i] = w[i] * J.dy_dv[i];
    for (int i = 0; i < Y; ++i) out[U + Y + i] = w[i] * J.dy_dn[i];
  } else {
    for (int i = 0; i < Y; ++i) out[U + i] = w[i] * J.dy_dn[i];
  }
}

static bool decompose_gram(const Model& M, const Jac& J, Gram& G) {
  const int U = M.U, Y = M.Y;
  for (int i = 0; i < Y; ++i) {
    double dv = M.family == FAMILY_HIER ? J.dy_dv[i] : 0.0;
    G.d[i] = dv * dv + J.dy_dn[i] * J.dy_dn[i];
  }
  if (M.dense) {
    for (int i = 0; i < Y; ++i)
      for (int j = 0; j < Y; ++j) {
        double s = (i == j) ? G.d[i] : 0.0;
        for (int a = 0; a < U; ++a)
          s += J.dy_du[(size_t)i * U + a] * J.dy_du[(size_t)j * U + a];
        G.chol[(size_t)i * Y + j] = s;
      }
    return chol_lower(G.chol.data(), Y);
  }
  for (int a = 0; a < U; ++a)
    for (int b = 0; b < U; ++b) {
      double s = (a == b) ? 1.0 : 0.0;
      for (int i = 0; i < Y; ++i)
        s += J.dy_du[(size_t)i * U + a] * (J.dy_du[(size_t)i * U + b] / G.d[i]);
      G.chol[(size_t)a * U + b] = s;
    }
  return chol_lower(G.chol.data(), U);
}

// out = Gram^-1 w  (tmp: >= U doubles)
static void lmult_inv_gram(const Model& M, const Jac& J, const Gram& G, const double* w,
                           double* out, double* tmp) {
  const int U = M.U, Y = M.Y;
  if (M.dense) {
    std::memcpy(out, w, sizeof(double) * Y);
    chol_solve(G.chol.data(), Y, out);
    return;
  }
  for (int a = 0; a < U; ++a) {
    double s = 0.0;
    for (int i = 0; i < Y; ++i) s += J.dy_du[(size_t)i * U + a] * (w[i] / G.d[i]);
    tmp[a] = s;
  }
  chol_solve(G.chol.data(), U, tmp);
  for (int i = 0; i < Y; ++i) {
    double s = 0.0;
    for (int a = 0; a < U; ++a) s += J.dy_du[(size_t)i * U + a] * tmp[a];
    out[i] = (w[i] - s) / G.d[i];
  }
}

// out = (J_l J_r^T)^-1 w   (mat: K*K scratch, tmp: >= max(U,Y))
static bool lmult_inv_jacob_product(const Model& M, const Jac& Jl, const Jac& Jr,
                                    const double* w, double* out, double* mat,
                                    double* tmp, double* dprod) {
  const int U = M.U, Y = M.Y;
  for (int i = 0; i < Y; ++i) {
    double dv = M.family == FAMILY_HIER ? Jl.dy_dv[i] * Jr.dy_dv[i] : 0.0;
    dprod[i] = dv + Jl.dy_dn[i] * Jr.dy_dn[i];
  }
  if (M.dense) {
    for (int i = 0; i < Y; ++i)
      for (int j = 0; j < Y; ++j) {
        double s = (i == j) ? dprod[i] : 0.0;
        for (int a = 0; a < U; ++a)
          s += Jl.dy_du[(size_t)i * U + a] * Jr.dy_du[(size_t)j * U + a];
        mat[(size_t)i * Y + j] = s;
      }
    std::memcpy(out, w, sizeof(double) * Y);
    return lu_solve(mat, Y, out);
  }
  for (int a = 0; a < U; ++a)
    for (int b = 0; b < U; ++b) {
      double s = (a == b) ? 1.0 : 0.0;
      for (int i = 0; i < Y; ++i)
        s += Jr.dy_du[(size_t)i * U + a] * (Jl.dy_du[(size_t)i * U + b] / dprod[i]);
      mat[(size_t)a * U + b] = s;
    }
  for (int a = 0; a < U; ++a) {
    double s = 0.0;
    for (int i = 0; i < Y; ++i) s += Jr.dy_du[(size_t)i * U + a] * (w[i] / dprod[i]);
    tmp[a] = s;
  }
  if (!lu_solve(mat, U, tmp)) return false;
  for (int i = 0; i < Y; ++i) {
    double s = 0.0;
    for (int a = 0; a < U; ++a) s += Jl.dy_du[(size_t)i * U + a] * tmp[a];
    out[i] = (w[i] - s) / dprod[i];
  }
  return true;
}

static double log_det_sqrt_gram(const Model& M, const Gram& G) {
  int K = M.dense ? M.Y : M.U;
  double s = 0.0;
  for (int k = 0; k < K; ++k) s += std::log(G.chol[(size_t)k * K + k]);
  if (!M.dense) {
    double t = 0.0;
    for (int i = 0; i < M.Y; ++i) t += std::log(G.d[i]);
    s += t / 2;
  }
  return s;
}

// ------------------------------------------------------------------- per-chain work

struct Work {
  Jac J, Jp, Jtmp;
  Gram G, Gp;
  std::vector<double> c, w, x, tq, tq2, mat, tmpu, dprod, mu, delta, q_try, dir;
  std::vector<double> mu_dt, qb, mub;
  // line search: set when a backtracking comparison |c(q + a d)| > |c(q)| was decided by
  // rounding (both norms below the residual noise floor, or equal to 1e-6 relative): such a
  // comparison, hence the halving count and the iteration path after it, is not
  // reproducible between two fp64 implementations (tests/test_gpu_full_size.py)
  int ls_ambiguous = 0;
  void init(const Model& M) {
    J.init(M), Jp.init(M), Jtmp.init(M), G.init(M), Gp.init(M);
    int K = M.dense ? M.Y : M.U;
    int mx = M.U > M.Y ? M.U : M.Y;
    c.assign(M.Y, 0), w.assign(M.Y, 0), x.assign(M.Y, 0);
    tq.assign(M.Q, 0), tq2.assign(M.Q, 0), mat.assign((size_t)K * K, 0);
    tmpu.assign(mx, 0), dprod.assign(M.Y, 0), mu.assign(M.Q, 0), delta.assign(M.Q, 0);
    q_try.assign(M.Q, 0), dir.assign(M.Q, 0);
    mu_dt.assign(M.Q, 0), qb.assign(M.Q, 0), mub.assign(M.Q, 0);
  }
};

static void lmult_pinv(const Model& M, Work& W, const Jac& J, const Gram& G,
                       const double* w, double* out) {
  lmult_inv_gram(M, J, G, w, W.x.data(), W.tmpu.data());
  rmult_jacob(M, J, W.x.data(), out);
}

static void normal_space_component(const Model& M, Work& W, const Jac& J, const Gram& G,
                                   const double* v, double* out) {
  lmult_jacob(M, J, v, W.w.data());
  lmult_pinv(M, W, J, G, W.w.data(), out);
}

enum { ST_OK = 0, ST_DIVERGED = 1, ST_NOT_CONVERGED = 2, ST_NON_REVERSIBLE = 3,
       ST_H_DIVERGED = 4, ST_LINALG = 5 };

enum { SOLVER_QUASI_NEWTON = 0, SOLVER_NEWTON = 1, SOLVER_NEWTON_LS = 2,
       SOLVER_MICI_NEWTON = 3, SOLVER_MICI_QUASI_NEWTON = 4 };

struct SolveOpts {
  double ctol, ptol, dtol;
  int max_iters, max_ls, norm;
};

struct SolveOut {
  int iters, n_constr, status;
  double ndq, err;
};

static int classify(double err, double ndq, const SolveOpts& o) {
  if (err < o.ctol && ndq < o.ptol) return ST_OK;
  if (err > o.dtol || std::isnan(err)) return ST_DIVERGED;
  return ST_NOT_CONVERGED;
}

static bool keep_iterating(int i, double ndq, double err, const SolveOpts& o) {
  bool diverged = (err > o.dtol) || std::isnan(err);
  bool converged = (err < o.ctol) && (ndq < o.ptol);
  return !((i >= o.max_iters) || diverged || converged);
}

// Project q (in/out) onto the manifold along the rows of Jp; mu_over_dt out.
static SolveOut solve_projection(const Model& M, Work& W, int solver, double* q,
                                 const Jac& Jp, const Gram& Gp, double dt,
                                 const SolveOpts& o, double* mu_over_dt) {
  const int Q = M.Q, Y = M.Y;
  SolveOut r{0, 0, ST_OK, INFINITY, -1.0};
  double* mu = W.mu.data();
  double* delta = W.delta.data();
  double* c = W.c.data();
  for (int k = 0; k < Q; ++k) mu[k] = 0.0;
  if (solver == SOLVER_QUASI_NEWTON || solver == SOLVER_NEWTON) {
    while (keep_iterating(r.iters, r.ndq, r.err, o)) {
      if (solver == SOLVER_QUASI_NEWTON) {
        M.constr(q, c);
        r.err = vec_norm(c, Y, o.norm);
        lmult_pinv(M, W, Jp, Gp, c, delta);
      } else {
        M.jacob(q, W.Jtmp.dy_du.data(), W.Jtmp.dy_dv.data(), W.Jtmp.dy_dn.data(), c);
        r.err = vec_norm(c, Y, o.norm);
        if (!lmult_inv_jacob_product(M, W.Jtmp, Jp, c, W.x.data(), W.mat.data(),
                                     W.tmpu.data(), W.dprod.data()))
          for (int i = 0; i < Y; ++i) W.x[i] = NAN;
        rmult_jacob(M, Jp, W.x.data(), delta);
      }
      for (int k = 0; k < Q; ++k) mu[k] += delta[k], q[k] -= delta[k];
      r.iters += 1;
      r.ndq = vec_norm(delta, Q, o.norm);
    }
    for (int k = 0; k < Q; ++k) mu_over_dt[k] = mu[k] / dt;
    r.status = classify(r.err, r.ndq, o);
    return r;
  }
  if (solver == SOLVER_NEWTON_LS) {
    double* q0 = W.tq2.data();
    double* dir = W.dir.data();
    double* qn = W.q_try.data();
    std::memcpy(q0, q, sizeof(double) * Q);
    while (keep_iterating(r.iters, r.ndq, r.err, o)) {
      M.jacob(q, W.Jtmp.dy_du.data(), W.Jtmp.dy_dv.data(), W.Jtmp.dy_dn.data(), c);
      double err_here = vec_norm(c, Y, o.norm);
      if (!lmult_inv_jacob_product(M, W.Jtmp, Jp, c, W.x.data(), W.mat.data(),
                                   W.tmpu.data(), W.dprod.data()))
        for (int i = 0; i < Y; ++i) W.x[i] = NAN;
      rmult_jacob(M, Jp, W.x.data(), dir);
      for (int k = 0; k < Q; ++k) dir[k] = -dir[k];
      double step = 1.0, new_err;
      int j = 0;
      do {
        for (int k = 0; k < Q; ++k) qn[k] = q[k] + step * dir[k];
        M.constr(qn, c);
        new_err = vec_norm(c, Y, o.norm);
        step *= 0.5;
        j += 1;
        {
          const double big = new_err > err_here ? new_err : err_here;
          if (big < 1e-10 || std::fabs(new_err - err_here) <= 1e-6 * big) W.ls_ambiguous = 1;
        }
      } while (j < o.max_ls && new_err > err_here);
      for (int k = 0; k < Q; ++k) delta[k] = qn[k] - q[k];
      r.ndq = vec_norm(delta, Q, o.norm);
      std::memcpy(q, qn, sizeof(double) * Q);
      r.err = new_err;
      r.iters += 1;
      r.n_constr += j;
    }
    for (int k = 0; k < Q; ++k) mu_over_dt[k] = (q0[k] - q[k]) / dt;
    r.status = classify(r.err, r.ndq, o);
    return r;
  }
  // Mici's own solvers: test BEFORE applying the update
  for (int i = 0; i < o.max_iters; ++i) {
    if (solver == SOLVER_MICI_NEWTON) {
      M.jacob(q, W.Jtmp.dy_du.data(), W.Jtmp.dy_dv.data(), W.Jtmp.dy_dn.data(), c);
      r.err = vec_norm(c, Y, o.norm);
      if (!lmult_inv_jacob_product(M, W.Jtmp, Jp, c, W.x.data(), W.mat.data(),
                                   W.tmpu.data(), W.dprod.data())) {
        r.status = ST_LINALG;
        return r;
      }
      rmult_jacob(M, Jp, W.x.data(), delta);
    } else {
      M.constr(q, c);
      r.err = vec_norm(c, Y, o.norm);
      lmult_pinv(M, W, Jp, Gp, c, delta);
    }
    // Mici: delta_mu = delta / |dt| and delta_pos = |dt| delta_mu = delta
    r.ndq = vec_norm(delta, Q, o.norm);
    r.iters = i;
    if (r.err > o.dtol || std::isnan(r.err)) {
      r.status = ST_DIVERGED;
      return r;
    }
    if (r.err < o.ctol && r.ndq < o.ptol) {
      for (int k = 0; k < Q; ++k) mu_over_dt[k] = mu[k] / dt;
      r.status = ST_OK;
      return r;
    }
    for (int k = 0; k < Q; ++k) mu[k] += delta[k], q[k] -= delta[k];
  }
  r.iters = o.max_iters;
  r.status = ST_NOT_CONVERGED;
  return r;
}

// ----------------------------------------------------------- leapfrog step / HMC

struct Chain {  // position, momentum and everything Mici caches at `q`
  std::vector<double> q, p, grad_h1;
  Jac J;
  Gram G;
  double nld, logdet;
  void init(const Model& M) {
    q.assign(M.Q, 0), p.assign(M.Q, 0), grad_h1.assign(M.Q, 0);
    J.init(M), G.init(M);
  }
};

// fill J, G, logdet, nld, grad_h1 at C.q ("dh1_dpos pre-evaluation")
static bool evaluate_at(const Model& M, Work& W, Chain& C) {
  M.jacob(C.q.data(), C.J.dy_du.data(), C.J.dy_dv.data(), C.J.dy_dn.data(), W.c.data());
  if (!decompose_gram(M, C.J, C.G)) return false;
  C.logdet = log_det_sqrt_gram(M, C.G);
  C.nld = M.nld(C.q.data(), C.grad_h1.data());
  M.grad_logdet(C.q.data(), C.J.dy_du.data(), C.J.dy_dv.data(), C.J.dy_dn.data(),
                C.G.chol.data(), C.G.d.data(), W.tq.data());
  for (int k = 0; k < M.Q; ++k) C.grad_h1[k] += W.tq[k];
  return true;
}

static void project_cotangent(const Model& M, Work& W, const Jac& J, const Gram& G,
                              double* p) {
  normal_space_component(M, W, J, G, p, W.tq.data());
  for (int k = 0; k < M.Q; ++k) p[k] -= W.tq[k];
}

static double hamiltonian(const Model& M, const Chain& C) {
  double s = 0.0;
  for (int k = 0; k < M.Q; ++k) s += C.p[k] * C.p[k];
  return C.nld + C.logdet + 0.5 * s;
}

struct StepOpts {
  int solver;
  SolveOpts so;
  double rev_tol;
  int rev_norm;
};

// One constrained leapfrog step A(dt/2) B(dt) A(dt/2) with reversibility check.
// `C` must be evaluated.  On failure C is left unchanged.
static int leapfrog_step(const Model& M, Work& W, Chain& C, Chain& N, double dt,
                         const StepOpts& o, int* it_fwd, int* it_rev) {
  const int Q = M.Q;
  N = C;
  // A
  for (int k = 0; k < Q; ++k) N.p[k] -= 0.5 * dt * C.grad_h1[k];
  project_cotangent(M, W, C.J, C.G, N.p.data());
  // B: position flow + retraction along rows of J(q_prev)
  for (int k = 0; k < Q; ++k) N.q[k] = C.q[k] + dt * N.p[k];
  std::vector<double>& mu_dt = W.mu_dt;
  SolveOut f = solve_projection(M, W, o.solver, N.q.data(), C.J, C.G, dt, o.so, mu_dt.data());
  *it_fwd = f.iters;
  *it_rev = 0;
  if (f.status != ST_OK) return f.status;
  for (int k = 0; k < Q; ++k) N.p[k] -= mu_dt[k];
  if (!evaluate_at(M, W, N)) return ST_LINALG;
  project_cotangent(M, W, N.J, N.G, N.p.data());
  // reverse check
  std::vector<double>&qb = W.qb, &mub = W.mub;
  for (int k = 0; k < Q; ++k) qb[k] = N.q[k] - dt * N.p[k];
  SolveOut b = solve_projection(M, W, o.solver, qb.data(), N.J, N.G, -dt, o.so, mub.data());
  *it_rev = b.iters;
  if (b.status != ST_OK) return b.status;
  for (int k = 0; k < Q; ++k) qb[k] -= C.q[k];
  if (vec_norm(qb.data(), Q, o.rev_norm) > o.rev_tol) return ST_NON_REVERSIBLE;
  // A
  for (int k = 0; k < Q; ++k) N.p[k] -= 0.5 * dt * N.grad_h1[k];
  project_cotangent(M, W, N.J, N.G, N.p.data());
  return ST_OK;
}

// ---------------------------------------------------------------- Philox4x32-10

static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
}

static inline double u01(uint32_t hi, uint32_t lo) {  // (0,1), 53 bits
  uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// stream 0: momentum normals (pair index `j` -> normals 2j, 2j+1); stream 1: uniform
static void philox_normal_pair(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t j,
                               double* n0, double* n1) {
  uint32_t c[4] = {j, iter, (uint32_t)chain, (uint32_t)(chain >> 32) & 0x7fffffffu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  double u1 = u01(c[0], c[1]), u2 = u01(c[2], c[3]);
  double r = std::sqrt(-2.0 * std::log(u1)), th = 6.283185307179586476925286766559 * u2;
  *n0 = r * std::cos(th), *n1 = r * std::sin(th);
}

static double philox_uniform(uint64_t seed, uint64_t chain, uint32_t iter) {
  uint32_t c[4] = {0u, iter, (uint32_t)chain,
                   ((uint32_t)(chain >> 32) & 0x7fffffffu) | 0x80000000u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01(c[0], c[1]);
}

// k-th uniform of one transition (k = 0 is philox_uniform: the Metropolis draw of the
// static sampler; the dynamic sampler numbers its draws 1, 2, ... in the order it uses them)
static double philox_uniform_k(uint64_t seed, uint64_t chain, uint32_t iter, uint32_t k) {
  uint32_t c[4] = {k, iter, (uint32_t)chain,
                   ((uint32_t)(chain >> 32) & 0x7fffffffu) | 0x80000000u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01(c[0], c[1]);
}

// ------------------------------------------------ dynamic multinomial HMC (NUTS)
// Restatement of Mici 0.2.1 `DynamicMultinomialHMC` (the sampler the reference builds at
// example_models/utils.py:286-288 with `max_tree_depth`, :70-73), from its published
// algorithm - Mici's source is not in the reference tree: recursive
// binary-tree doubling in a uniformly drawn direction, multinomial sampling of the
// proposal with weights exp(h_init - h) (uniform sampling inside a new subtree, biased
// progressive sampling between the old tree and the new subtree), no-U-turn termination
// on the summed momentum (identity metric: dh/dmom = p) checked for every subtree and for
// the whole tree and - `do_extra_subtree_checks`, Mici's default - for the two overlapping
// sub-trajectories of every merge (the subtree built first plus the first state of the one
// built after it; the last state of the first plus the second), termination on an
// integrator error or h - h_init > max_delta_h (one-sided, as Mici's dynamic transitions
// test it).  accept_stat = mean over the tree's integrator steps of min(1, exp(h_init - h))
// (a failed step counts as 0).  NUTS_NO_EXTRA_CHECKS / NUTS_TWO_SIDED (flags of
// orc_solver_opts, same values as include/mlb.h) switch the two refinements off.
struct Tree {
  Chain neg, pos;            // edge states (time order), evaluated
  std::vector<double> rho;   // sum of momenta over the tree's states
  std::vector<double> prop;  // proposal position
  double lw;                 // log of the summed weights
  void init(const Model& M) { neg.init(M), pos.init(M), rho.assign(M.Q, 0), prop.assign(M.Q, 0); }
};

struct NutsCtx {
  const Model& M;
  Work& W;
  const StepOpts& so;
  double dt, h0, max_delta_h;
  bool extra_checks, two_sided;
  uint64_t seed, chain;
  uint32_t iter, k;  // next uniform index
  const double* unif;  // optional supplied draws (index k - 1)
  int n_unif;
  int n_step, status;  // integrator steps taken; first failure status
  double sum_acc;
  double draw() {
    double u = (unif && (int)k - 1 < n_unif) ? unif[k - 1] : philox_uniform_k(seed, chain, iter, k);
    ++k;
    return u;
  }
};

static double log_add_exp(double a, double b) {
  double m = a > b ? a : b;
  return m + std::log(std::exp(a - m) + std::exp(b - m));
}

static bool no_u_turn_violated(const Model& M, const Chain& a, const Chain& b,
                               const std::vector<double>& rho) {
  double da = 0.0, db = 0.0;
  for (int k = 0; k < M.Q; ++k) da += a.p[k] * rho[k], db += b.p[k] * rho[k];
  return da < 0.0 || db < 0.0;
}

// The criterion on the two overlapping sub-trajectories when `outer` (built in direction
// dir after `inner`) joins `inner`: inner + first state of outer, last state of inner + outer.
static bool extra_checks_violated(const Model& M, const Tree& inner, const Tree& outer, int dir) {
  const Chain& in_first = dir > 0 ? inner.neg : inner.pos;
  const Chain& in_last = dir > 0 ? inner.pos : inner.neg;
  const Chain& out_first = dir > 0 ? outer.neg : outer.pos;
  const Chain& out_last = dir > 0 ? outer.pos : outer.neg;
  std::vector<double> rho(M.Q);
  for (int k = 0; k < M.Q; ++k) rho[k] = inner.rho[k] + out_first.p[k];
  if (no_u_turn_violated(M, in_first, out_first, rho)) return true;
  for (int k = 0; k < M.Q; ++k) rho[k] = outer.rho[k] + in_last.p[k];
  return no_u_turn_violated(M, in_last, out_last, rho);
}

// Builds a subtree of 2^depth states beyond `edge` in direction dir; returns true if the
// transition must terminate (the subtree is then discarded).
static bool build_tree(NutsCtx& X, int depth, const Chain& edge, int dir, Tree& out) {
  const Model& M = X.M;
  if (depth == 0) {
    Chain N;
    N.init(M);
    int f, r;
    int st = leapfrog_step(M, X.W, const_cast<Chain&>(edge), N, dir * X.dt, X.so, &f, &r);
    X.n_step += 1;
    if (st != ST_OK) {
      if (X.status == ST_OK) X.status = st;
      return true;
    }
    double h = hamiltonian(M, N);
    if (std::isnan(h) || h - X.h0 > X.max_delta_h ||
        (X.two_sided && X.h0 - h > X.max_delta_h)) {
      if (X.status == ST_OK) X.status = ST_H_DIVERGED;
      return true;
    }
    double d = X.h0 - h;
    X.sum_acc += std::exp(d < 0.0 ? d : 0.0);
    out.neg = N, out.pos = N, out.rho = N.p, out.prop = N.q, out.lw = d;
    return false;
  }
  Tree inner, outer;
  inner.init(M), outer.init(M);
  if (build_tree(X, depth - 1, edge, dir, inner)) return true;
  if (build_tree(X, depth - 1, dir > 0 ? inner.pos : inner.neg, dir, outer)) return true;
  const double lw = log_add_exp(inner.lw, outer.lw);
  const double u = X.draw();
  out.prop = (u < std::exp(outer.lw - lw)) ? outer.prop : inner.prop;
  out.lw = lw;
  out.rho = inner.rho;
  for (int k = 0; k < M.Q; ++k) out.rho[k] += outer.rho[k];
  if (dir > 0) out.neg = inner.neg, out.pos = outer.pos;
  else out.neg = outer.neg, out.pos = inner.pos;
  if (no_u_turn_violated(M, out.neg, out.pos, out.rho)) return true;
  return X.extra_checks && extra_checks_violated(M, inner, outer, dir);
}

}  // namespace orc

// ======================================================================== C API
using namespace orc;

extern "C" {

typedef struct orc_solver_opts {
  int solver, max_iters, max_ls, norm;
  double ctol, ptol, dtol;
  double rev_tol;
  int rev_norm, flags;  // flags: NUTS_* below (values of include/mlb.h MLB_FLAG_NUTS_*)
} orc_solver_opts;
enum { NUTS_NO_EXTRA_CHECKS = 2, NUTS_TWO_SIDED = 4 };

static StepOpts to_step_opts(const orc_solver_opts* o) {
  StepOpts s;
  s.solver = o->solver;
  s.so = SolveOpts{o->ctol, o->ptol, o->dtol, o->max_iters, o->max_ls, o->norm};
  s.rev_tol = o->rev_tol;
  s.rev_norm = o->rev_norm;
  return s;
}

void* orc_model_create(int model_id, int dim_y, int param, const double* data,
                       int n_data) {
  return (void*)make_model(model_id, dim_y, param, data, n_data);
}
void orc_model_destroy(void* h) { delete (Model*)h; }
int orc_dim_u(void* h) { return ((Model*)h)->U; }
int orc_dim_y(void* h) { return ((Model*)h)->Y; }
int orc_dim_q(void* h) { return ((Model*)h)->Q; }
int orc_is_dense(void* h) { return ((Model*)h)->dense ? 1 : 0; }
// torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU legs set the thread
// count explicitly so that they use the host cores whatever launched them
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static std::vector<int32_t> g_ls_ambiguous;  // per chain, of the last solve / step call

#define ORC_FOR_CHAINS(body)                                   \
  _Pragma("omp parallel") {                                    \
    Work W;                                                    \
    W.init(M);                                                 \
    _Pragma("omp for schedule(static)") for (int64_t i = 0; i < n; ++i) { body } \
  }

void orc_constr(void* h, int64_t n, const double* q, double* c) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(M.constr(q + i * M.Q, c + i * M.Y);)
}

// blocks out: dy_du [n][Y][U], dy_dv [n][Y] (hier only, may be null), dy_dn [n][Y]
void orc_jacob_constr_blocks(void* h, int64_t n, const double* q, double* dy_du,
                             double* dy_dv, double* dy_dn, double* c) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(
      M.jacob(q + i * M.Q, dy_du + i * M.Y * M.U, dy_dv ? dy_dv + i * M.Y : nullptr,
              dy_dn + i * M.Y, c + i * M.Y);)
}

static void load_jac(const Model& M, Jac& J, int64_t i, const double* dy_du,
                     const double* dy_dv, const double* dy_dn) {
  std::memcpy(J.dy_du.data(), dy_du + i * M.Y * M.U, sizeof(double) * M.Y * M.U);
  if (M.family == FAMILY_HIER)
    std::memcpy(J.dy_dv.data(), dy_dv + i * M.Y, sizeof(double) * M.Y);
  std::memcpy(J.dy_dn.data(), dy_dn + i * M.Y, sizeof(double) * M.Y);
}

// gram out: chol [n][K][K] (K = Y dense, U Woodbury), d [n][Y]
void orc_decompose_gram(void* h, int64_t n, const double* dy_du, const double* dy_dv,
                        const double* dy_dn, double* chol, double* d) {
  const Model& M = *(Model*)h;
  const int K = M.dense ? M.Y : M.U;
  ORC_FOR_CHAINS(load_jac(M, W.J, i, dy_du, dy_dv, dy_dn); decompose_gram(M, W.J, W.G);
                 std::memcpy(chol + i * K * K, W.G.chol.data(), sizeof(double) * K * K);
                 std::memcpy(d + i * M.Y, W.G.d.data(), sizeof(double) * M.Y);)
}

void orc_log_det_sqrt_gram(void* h, int64_t n, const double* q, double* val) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(M.jacob(q + i * M.Q, W.J.dy_du.data(), W.J.dy_dv.data(),
                         W.J.dy_dn.data(), W.c.data());
                 decompose_gram(M, W.J, W.G); val[i] = log_det_sqrt_gram(M, W.G);)
}

void orc_grad_log_det_sqrt_gram(void* h, int64_t n, const double* q, double* grad,
                                double* val) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(M.jacob(q + i * M.Q, W.J.dy_du.data(), W.J.dy_dv.data(),
                         W.J.dy_dn.data(), W.c.data());
                 decompose_gram(M, W.J, W.G); val[i] = log_det_sqrt_gram(M, W.G);
                 M.grad_logdet(q + i * M.Q, W.J.dy_du.data(), W.J.dy_dv.data(),
                               W.J.dy_dn.data(), W.G.chol.data(), W.G.d.data(),
                               grad + i * M.Q);)
}

void orc_neg_log_dens(void* h, int64_t n, const double* q, double* val, double* grad) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(val[i] = M.nld(q + i * M.Q, grad ? grad + i * M.Q : nullptr);)
}

void orc_lmult_by_jacob_constr(void* h, int64_t n, const double* dy_du,
                               const double* dy_dv, const double* dy_dn, const double* v,
                               double* out) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(load_jac(M, W.J, i, dy_du, dy_dv, dy_dn);
                 lmult_jacob(M, W.J, v + i * M.Q, out + i * M.Y);)
}

void orc_rmult_by_jacob_constr(void* h, int64_t n, const double* dy_du,
                               const double* dy_dv, const double* dy_dn, const double* w,
                               double* out) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(load_jac(M, W.J, i, dy_du, dy_dv, dy_dn);
                 rmult_jacob(M, W.J, w + i * M.Y, out + i * M.Q);)
}

static void load_gram(const Model& M, Gram& G, int64_t i, const double* chol,
                      const double* d) {
  const int K = M.dense ? M.Y : M.U;
  std::memcpy(G.chol.data(), chol + i * K * K, sizeof(double) * K * K);
  std::memcpy(G.d.data(), d + i * M.Y, sizeof(double) * M.Y);
}

void orc_lmult_by_pinv_jacob_constr(void* h, int64_t n, const double* dy_du,
                                    const double* dy_dv, const double* dy_dn,
                                    const double* chol, const double* d, const double* w,
                                    double* out) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(load_jac(M, W.J, i, dy_du, dy_dv, dy_dn); load_gram(M, W.G, i, chol, d);
                 lmult_pinv(M, W, W.J, W.G, w + i * M.Y, out + i * M.Q);)
}

void orc_normal_space_component(void* h, int64_t n, const double* dy_du,
                                const double* dy_dv, const double* dy_dn,
                                const double* chol, const double* d, const double* v,
                                double* out) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(load_jac(M, W.J, i, dy_du, dy_dv, dy_dn); load_gram(M, W.G, i, chol, d);
                 normal_space_component(M, W, W.J, W.G, v + i * M.Q, out + i * M.Q);)
}

void orc_lmult_by_inv_jacob_product(void* h, int64_t n, const double* dy_du_l,
                                    const double* dy_dv_l, const double* dy_dn_l,
                                    const double* dy_du_r, const double* dy_dv_r,
                                    const double* dy_dn_r, const double* w, double* out) {
  const Model& M = *(Model*)h;
  ORC_FOR_CHAINS(load_jac(M, W.J, i, dy_du_l, dy_dv_l, dy_dn_l);
                 load_jac(M, W.Jp, i, dy_du_r, dy_dv_r, dy_dn_r);
                 lmult_inv_jacob_product(M, W.J, W.Jp, w + i * M.Y, out + i * M.Y,
                                         W.mat.data(), W.tmpu.data(), W.dprod.data());)
}

// Project q[i] onto the manifold using the Jacobian / Gram evaluated at q_prev[i].
void orc_solve_projection(void* h, int64_t n, const double* q, const double* q_prev,
                          double dt, const orc_solver_opts* opts, double* q_out,
                          double* mu_out, int32_t* iters, double* norm_dq, double* err,
                          int32_t* n_constr, int32_t* status) {
  const Model& M = *(Model*)h;
  StepOpts so = to_step_opts(opts);
  g_ls_ambiguous.assign((size_t)n, 0);
  ORC_FOR_CHAINS(
      W.ls_ambiguous = 0;
      M.jacob(q_prev + i * M.Q, W.Jp.dy_du.data(), W.Jp.dy_dv.data(), W.Jp.dy_dn.data(),
              W.c.data());
      decompose_gram(M, W.Jp, W.Gp);
      std::memcpy(q_out + i * M.Q, q + i * M.Q, sizeof(double) * M.Q);
      SolveOut r = solve_projection(M, W, so.solver, q_out + i * M.Q, W.Jp, W.Gp, dt,
                                    so.so, mu_out + i * M.Q);
      iters[i] = r.iters; norm_dq[i] = r.ndq; err[i] = r.err; n_constr[i] = r.n_constr;
      status[i] = r.status; g_ls_ambiguous[(size_t)i] = W.ls_ambiguous;)
}

// n_step constrained leapfrog steps per chain.  dt_per_chain may be null (use dt).
// Outputs: q,p updated in place to the last successfully completed step; status of the
// first failing step (or OK); n_done steps completed; iteration totals; h at start/end.
void orc_leapfrog_steps(void* h, int64_t n, double* q, double* p, double dt,
                        const double* dt_per_chain, int n_step,
                        const orc_solver_opts* opts, int32_t* status, int32_t* n_done,
                        int32_t* iters_fwd, int32_t* iters_rev, double* h_init,
                        double* h_final) {
  const Model& M = *(Model*)h;
  StepOpts so = to_step_opts(opts);
  g_ls_ambiguous.assign((size_t)n, 0);
#pragma omp parallel
  {
    Work W;
    W.init(M);
    Chain C, N;
    C.init(M), N.init(M);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      W.ls_ambiguous = 0;
      std::memcpy(C.q.data(), q + i * M.Q, sizeof(double) * M.Q);
      std::memcpy(C.p.data(), p + i * M.Q, sizeof(double) * M.Q);
      double dti = dt_per_chain ? dt_per_chain[i] : dt;
      int st = evaluate_at(M, W, C) ? ST_OK : ST_LINALG;
      if (h_init) h_init[i] = hamiltonian(M, C);
      int done = 0, tf = 0, tr = 0;
      for (int s = 0; s < n_step && st == ST_OK; ++s) {
        int f, r;
        st = leapfrog_step(M, W, C, N, dti, so, &f, &r);
        tf += f, tr += r;
        if (st == ST_OK) std::swap(C, N), done += 1;
      }
      std::memcpy(q + i * M.Q, C.q.data(), sizeof(double) * M.Q);
      std::memcpy(p + i * M.Q, C.p.data(), sizeof(double) * M.Q);
      status[i] = st, n_done[i] = done;
      g_ls_ambiguous[(size_t)i] = W.ls_ambiguous;
      if (iters_fwd) iters_fwd[i] = tf;
      if (iters_rev) iters_rev[i] = tr;
      if (h_final) h_final[i] = hamiltonian(M, C);
    }
  }
}

// One static-HMC transition per chain (notebook `simulate_proposal`): momentum
// refresh (projected N(0,I)), n_step leapfrog steps with |dh| > max_delta_h divergence
// test after every step, Metropolis accept.  Random numbers: supplied (`mom_noise`
// [n][Q], `unif` [n]) or, when null, Philox4x32-10 keyed by (seed, chain0 + i, iter).
void orc_hmc_transition(void* h, int64_t n, double* q, double dt,
                        const double* dt_per_chain, int n_step,
                        const orc_solver_opts* opts, double max_delta_h,
                        const double* mom_noise, const double* unif, uint64_t seed,
                        uint64_t chain0, uint32_t iter, int32_t* status,
                        int32_t* accepted, double* accept_stat, int32_t* n_done,
                        int32_t* iters_total, double* h_out) {
  const Model& M = *(Model*)h;
  StepOpts so = to_step_opts(opts);
#pragma omp parallel
  {
    Work W;
    W.init(M);
    Chain C, N, S;
    C.init(M), N.init(M), S.init(M);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const int Q = M.Q;
      std::memcpy(C.q.data(), q + i * Q, sizeof(double) * Q);
      if (mom_noise) {
        std::memcpy(C.p.data(), mom_noise + i * Q, sizeof(double) * Q);
      } else {
        for (int j = 0; 2 * j < Q; ++j) {
          double a, b;
          philox_normal_pair(seed, chain0 + i, iter, (uint32_t)j, &a, &b);
          C.p[2 * j] = a;
          if (2 * j + 1 < Q) C.p[2 * j + 1] = b;
        }
      }
      double u = unif ? unif[i] : philox_uniform(seed, chain0 + i, iter);
      double dti = dt_per_chain ? dt_per_chain[i] : dt;
      int st = evaluate_at(M, W, C) ? ST_OK : ST_LINALG;
      project_cotangent(M, W, C.J, C.G, C.p.data());
      S = C;
      double h0 = hamiltonian(M, C), h1 = h0;
      int done = 0, tot = 0;
      for (int s = 0; s < n_step && st == ST_OK; ++s) {
        int f, r;
        st = leapfrog_step(M, W, C, N, dti, so, &f, &r);
        tot += f + r;
        if (st == ST_OK) {
          std::swap(C, N), done += 1;
          h1 = hamiltonian(M, C);
          if (std::fabs(h1 - h0) > max_delta_h) st = ST_H_DIVERGED;
        }
      }
      double a_stat = 0.0;
      int acc = 0;
      if (st == ST_OK) {
        double d = h0 - h1;
        a_stat = std::isnan(d) ? 0.0 : std::exp(d < 0.0 ? d : 0.0);
        acc = u < a_stat ? 1 : 0;
      }
      const Chain& R = acc ? C : S;
      std::memcpy(q + i * Q, R.q.data(), sizeof(double) * Q);
      status[i] = st, accepted[i] = acc, accept_stat[i] = a_stat, n_done[i] = done;
      if (iters_total) iters_total[i] = tot;
      if (h_out) h_out[i] = acc ? h1 : h0;
    }
  }
}

// One dynamic multinomial HMC transition per chain.  unif: optional [n][n_unif] supplied
// uniforms (draw k of the transition reads unif[i][k - 1]; Philox beyond n_unif).
void orc_nuts_transition(void* h, int64_t n, double* q, double dt,
                         const double* dt_per_chain, int max_depth,
                         const orc_solver_opts* opts, double max_delta_h,
                         const double* mom_noise, const double* unif, int n_unif,
                         uint64_t seed, uint64_t chain0, uint32_t iter, int32_t* status,
                         double* accept_stat, int32_t* n_step, int32_t* depth_out) {
  const Model& M = *(Model*)h;
  StepOpts so = to_step_opts(opts);
#pragma omp parallel
  {
    Work W;
    W.init(M);
    Chain C;
    C.init(M);
    Tree T, S;
    T.init(M), S.init(M);
#pragma omp for schedule(dynamic, 4)
    for (int64_t i = 0; i < n; ++i) {
      const int Q = M.Q;
      std::memcpy(C.q.data(), q + i * Q, sizeof(double) * Q);
      if (mom_noise) {
        std::memcpy(C.p.data(), mom_noise + i * Q, sizeof(double) * Q);
      } else {
        for (int j = 0; 2 * j < Q; ++j) {
          double a, b;
          philox_normal_pair(seed, chain0 + i, iter, (uint32_t)j, &a, &b);
          C.p[2 * j] = a;
          if (2 * j + 1 < Q) C.p[2 * j + 1] = b;
        }
      }
      int st0 = evaluate_at(M, W, C) ? ST_OK : ST_LINALG;
      project_cotangent(M, W, C.J, C.G, C.p.data());
      NutsCtx X{M, W, so, dt_per_chain ? dt_per_chain[i] : dt, hamiltonian(M, C), max_delta_h,
                !(opts->flags & NUTS_NO_EXTRA_CHECKS), (opts->flags & NUTS_TWO_SIDED) != 0, seed, (uint64_t)(chain0 + i), iter, 1u, unif ? unif + i * n_unif : nullptr,
                n_unif, 0, st0, 0.0};
      T.neg = C, T.pos = C, T.rho = C.p, T.prop = C.q, T.lw = 0.0;
      int depth = 0;
      for (; depth < max_depth && st0 == ST_OK; ++depth) {
        const int dir = X.draw() < 0.5 ? 1 : -1;
        if (build_tree(X, depth, dir > 0 ? T.pos : T.neg, dir, S)) break;
        if (X.draw() < std::exp(S.lw - T.lw)) T.prop = S.prop;
        T.lw = log_add_exp(T.lw, S.lw);
        const bool extra_turn = X.extra_checks && extra_checks_violated(M, T, S, dir);
        for (int k = 0; k < Q; ++k) T.rho[k] += S.rho[k];
        if (dir > 0) T.pos = S.pos; else T.neg = S.neg;
        if (extra_turn || no_u_turn_violated(M, T.neg, T.pos, T.rho)) {
          ++depth;
          break;
        }
      }
      std::memcpy(q + i * Q, T.prop.data(), sizeof(double) * Q);
      status[i] = X.status;
      accept_stat[i] = X.n_step > 0 ? X.sum_acc / X.n_step : 0.0;
      n_step[i] = X.n_step;
      if (depth_out) depth_out[i] = depth;
    }
  }
}

// per-chain flags of the last orc_solve_projection / orc_leapfrog_steps call (see Work)
void orc_last_line_search_ambiguous(int32_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = (size_t)i < g_ls_ambiguous.size() ? g_ls_ambiguous[(size_t)i] : 0;
}

void orc_philox_normals(uint64_t seed, uint64_t chain, uint32_t iter, int n, double* out) {
  for (int j = 0; 2 * j < n; ++j) {
    double a, b;
    philox_normal_pair(seed, chain, iter, (uint32_t)j, &a, &b);
    out[2 * j] = a;
    if (2 * j + 1 < n) out[2 * j + 1] = b;
  }
}
double orc_philox_uniform(uint64_t seed, uint64_t chain, uint32_t iter) {
  return philox_uniform(seed, chain, iter);
}

}  // extern "C"
