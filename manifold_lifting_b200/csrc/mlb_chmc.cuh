// Generic per-thread constrained-HMC device code: projection solvers and the
// constrained leapfrog step, templated on a model (see mlb_models.cuh).
//
// Restates, for one chain held in one thread's registers:
//   * the three projection loops of mlift/systems.py:190-228 (quasi-Newton),
//     230-270 (Newton), 272-326 (Newton + backtracking line search) with the same
//     carried values, initial values (q, 0, 0, inf, -1) and continuation predicate,
//   * the success / divergence classification of mlift/solvers.py:81-95,
//   * Mici's own Newton / quasi-Newton solvers (test-before-update), used by the toy
//     configs (toy_2d_model.py:170-174),
//   * Mici's ConstrainedLeapfrogIntegrator step A(dt/2) B(dt) A(dt/2) with the
//     reversibility check (call site utils.py:322-327).
// Per-chain failure is a status code, never an exception.
#pragma once
#include "mlb_models.cuh"

namespace mlb {

struct SolveOut {
  int iters, n_constr, status;
  double ndq, err;
};

MLB_DEV bool keep_iterating(int i, double ndq, double err, const Opts& o) {
  bool diverged = (err > o.dtol) || (err != err);
  bool converged = (err < o.ctol) && (ndq < o.ptol);
  return !((i >= o.max_iters) || diverged || converged);
}

MLB_DEV int classify(double err, double ndq, const Opts& o) {
  if (err < o.ctol && ndq < o.ptol) return MLB_ST_OK;
  if (err > o.dtol || err != err) return MLB_ST_DIVERGED;
  return MLB_ST_NOT_CONVERGED;
}

// Project q (in/out) onto {c = 0} along the rows of Jp.  mu_dt <- (sum of updates)/dt,
// i.e. what the reference subtracts from the momentum (solvers.py:84).
template <class M, int SOLVER>
MLB_DEV SolveOut solve_projection(const typename M::Params& P, double (&q)[M::Q],
                                  const typename M::B::Jac& Jp,
                                  const typename M::B::Gram& Gp, double dt, const Opts& o,
                                  double (&mu_dt)[M::Q]) {
  using B = typename M::B;
  constexpr int Q = M::Q, Y = M::Y;
  SolveOut r;
  r.iters = 0, r.n_constr = 0, r.status = MLB_ST_OK, r.ndq = INFINITY, r.err = -1.0;
  const double rdt = 1.0 / dt;

  if (SOLVER == MLB_SOLVER_QUASI_NEWTON || SOLVER == MLB_SOLVER_NEWTON) {
    double mu[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) mu[k] = 0.0;
    while (keep_iterating(r.iters, r.ndq, r.err, o)) {
      double c[Y], delta[Q];
      if (SOLVER == MLB_SOLVER_QUASI_NEWTON) {
        M::constr(P, q, c);
        r.err = vec_norm<Y>(c, o.norm);
        B::lmult_pinv(Jp, Gp, c, delta);
      } else {
        typename B::Jac Jc;
        M::jacob(P, q, Jc, c);
        r.err = vec_norm<Y>(c, o.norm);
        double x[Y];
        bool ok = B::lmult_inv_jacob_product(Jc, Jp, c, x);
        if (!ok) {
#pragma unroll
          for (int i = 0; i < Y; ++i) x[i] = nan("");
        }
        B::rmult_jacob(Jp, x, delta);
      }
#pragma unroll
      for (int k = 0; k < Q; ++k) mu[k] += delta[k], q[k] -= delta[k];
      r.iters += 1;
      r.ndq = vec_norm<Q>(delta, o.norm);
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) mu_dt[k] = mu[k] / dt;
    r.status = classify(r.err, r.ndq, o);
    return r;
  }

  if (SOLVER == MLB_SOLVER_NEWTON_LS) {
    double q0[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) q0[k] = q[k];
    while (keep_iterating(r.iters, r.ndq, r.err, o)) {
      double c[Y], dir[Q], x[Y], qn[Q];
      typename B::Jac Jc;
      M::jacob(P, q, Jc, c);
      const double err_here = vec_norm<Y>(c, o.norm);
      bool ok = B::lmult_inv_jacob_product(Jc, Jp, c, x);
      if (!ok) {
#pragma unroll
        for (int i = 0; i < Y; ++i) x[i] = nan("");
      }
      B::rmult_jacob(Jp, x, dir);
      double step = 1.0, new_err;
      int j = 0;
      do {  // first trial unconditional, then halve while the residual grew
#pragma unroll
        for (int k = 0; k < Q; ++k) qn[k] = fma(-step, dir[k], q[k]);
        M::constr(P, qn, c);
        new_err = vec_norm<Y>(c, o.norm);
        step *= 0.5;
        j += 1;
      } while (j < o.max_ls && new_err > err_here);
      double dq[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) dq[k] = qn[k] - q[k], q[k] = qn[k];
      r.ndq = vec_norm<Q>(dq, o.norm);
      r.err = new_err;
      r.iters += 1;
      r.n_constr += j;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) mu_dt[k] = (q0[k] - q[k]) / dt;
    r.status = classify(r.err, r.ndq, o);
    return r;
  }

  // Mici's own solvers: convergence tested BEFORE the update is applied.
  {
    double mu[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) mu[k] = 0.0;
    r.status = MLB_ST_NOT_CONVERGED;
    int i = 0;
    for (; i < o.max_iters; ++i) {
      double c[Y], delta[Q];
      if (SOLVER == MLB_SOLVER_MICI_NEWTON) {
        typename B::Jac Jc;
        M::jacob(P, q, Jc, c);
        r.err = vec_norm<Y>(c, o.norm);
        double x[Y];
        if (!B::lmult_inv_jacob_product(Jc, Jp, c, x)) {
          r.status = MLB_ST_LINALG;
          break;
        }
        B::rmult_jacob(Jp, x, delta);
      } else {
        M::constr(P, q, c);
        r.err = vec_norm<Y>(c, o.norm);
        B::lmult_pinv(Jp, Gp, c, delta);
      }
      r.ndq = vec_norm<Q>(delta, o.norm);
      if (r.err > o.dtol || r.err != r.err) {
        r.status = MLB_ST_DIVERGED;
        break;
      }
      if (r.err < o.ctol && r.ndq < o.ptol) {
        r.status = MLB_ST_OK;
        break;
      }
#pragma unroll
      for (int k = 0; k < Q; ++k) mu[k] += delta[k], q[k] -= delta[k];
    }
    r.iters = i;
#pragma unroll
    for (int k = 0; k < Q; ++k) mu_dt[k] = mu[k] * rdt;
    return r;
  }
}

// Everything Mici caches at a position (systems.py:369-405): Jacobian blocks, Gram
// components, log det term, prior nld, and dh1_dpos = grad nld + grad 0.5 log det Gram.
template <class M>
struct Eval {
  typename M::B::Jac J;
  typename M::B::Gram G;
  double g[M::Q];
  double nld, logdet;
};

template <class M>
MLB_DEV bool evaluate_at(const typename M::Params& P, const double (&q)[M::Q], Eval<M>& E) {
  double c[M::Y];
  M::jacob(P, q, E.J, c);
  M::B::decompose_gram(E.J, E.G);
  E.logdet = M::B::log_det_sqrt_gram(E.G);
  E.nld = M::nld(P, q, E.g);
  M::add_grad_logdet(P, q, E.J, E.G, E.g);
  return E.G.ok;
}

template <int Q>
MLB_DEV double half_sq_norm(const double (&p)[Q]) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < Q; ++k) s = fma(p[k], p[k], s);
  return 0.5 * s;
}

// One constrained leapfrog step.  On success q, p, E are advanced and MLB_ST_OK is
// returned; on failure they are left at the pre-step values.
template <class M, int SOLVER>
MLB_DEV int leapfrog_step(const typename M::Params& P, double (&q)[M::Q], double (&p)[M::Q],
                          Eval<M>& E, double dt, const Opts& o, int& it_fwd, int& it_rev) {
  using B = typename M::B;
  constexpr int Q = M::Q;
  it_fwd = 0, it_rev = 0;
  double pn[Q], qn[Q];
  // A(dt/2): h1 flow + cotangent projection at q
#pragma unroll
  for (int k = 0; k < Q; ++k) pn[k] = fma(-0.5 * dt, E.g[k], p[k]);
  B::project_cotangent(E.J, E.G, pn);
  // B(dt): h2 flow, retraction onto the manifold along rows of J(q)
#pragma unroll
  for (int k = 0; k < Q; ++k) qn[k] = fma(dt, pn[k], q[k]);
  {
    double mu_dt[Q];
    SolveOut f = solve_projection<M, SOLVER>(P, qn, E.J, E.G, dt, o, mu_dt);
    it_fwd = f.iters;
    if (f.status != MLB_ST_OK) return f.status;
#pragma unroll
    for (int k = 0; k < Q; ++k) pn[k] -= mu_dt[k];
  }
  Eval<M> En;
  if (!evaluate_at<M>(P, qn, En)) return MLB_ST_LINALG;
  B::project_cotangent(En.J, En.G, pn);
  {  // reversibility check: step back from (qn, pn) and compare with q
    double qb[Q], mub[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) qb[k] = fma(-dt, pn[k], qn[k]);
    SolveOut b = solve_projection<M, SOLVER>(P, qb, En.J, En.G, -dt, o, mub);
    it_rev = b.iters;
    if (b.status != MLB_ST_OK) return b.status;
#pragma unroll
    for (int k = 0; k < Q; ++k) qb[k] -= q[k];
    if (vec_norm<Q>(qb, o.rev_norm) > o.rev_tol) return MLB_ST_NON_REVERSIBLE;
  }
  // A(dt/2) at the new position
#pragma unroll
  for (int k = 0; k < Q; ++k) pn[k] = fma(-0.5 * dt, En.g[k], pn[k]);
  B::project_cotangent(En.J, En.G, pn);
#pragma unroll
  for (int k = 0; k < Q; ++k) q[k] = qn[k], p[k] = pn[k];
  E = En;
  return MLB_ST_OK;
}

}  // namespace mlb
