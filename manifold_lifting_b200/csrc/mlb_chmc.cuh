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

#ifndef MLB_FUSE_KICKS
#define MLB_FUSE_KICKS 1  // one kick + projection at interior step boundaries (see below)
#endif
#ifndef MLB_HI_NORM
#define MLB_HI_NORM 1  // high-word maximum norm in the fused solver loops (0: exact always)
#endif

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
// WANT_MU = false: mu_dt is not produced (the fused trajectory recovers the momentum
// correction from the positions, mu = q_start - q_end, without carrying Q accumulators).
template <class M, int SOLVER, bool WANT_MU = true>
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
    // the fused drivers (WANT_MU = false) only compare the norms with the tolerances:
    // under the maximum norm they use the high-word norm (mlb_common.cuh) and fall back to
    // the exact one when a comparison would be ambiguous
    const bool hi_norm = MLB_HI_NORM && !WANT_MU && o.norm == MLB_NORM_MAX;
    auto residual_norm = [&](const double (&c)[Y]) {
      if (!hi_norm) return vec_norm<Y>(c, o.norm);
      const double t = max_norm_hi<Y>(c);
      return (norm_hi_ambiguous(t, o.ctol) || norm_hi_ambiguous(t, o.dtol))
                 ? vec_norm<Y>(c, MLB_NORM_MAX) : t;
    };
    auto update_norm = [&](const double (&d)[Q]) {
      if (!hi_norm) return vec_norm<Q>(d, o.norm);
      const double t = max_norm_hi<Q>(d);
      return norm_hi_ambiguous(t, o.ptol) ? vec_norm<Q>(d, MLB_NORM_MAX) : t;
    };
    while (keep_iterating(r.iters, r.ndq, r.err, o)) {
      double c[Y], delta[Q];
      if (SOLVER == MLB_SOLVER_QUASI_NEWTON) {
        M::constr(P, q, c);
        r.err = residual_norm(c);
        B::lmult_pinv(Jp, Gp, c, delta);
      } else {
        typename B::Jac Jc;
        M::jacob(P, q, Jc, c);
        r.err = residual_norm(c);
        double x[Y];
        bool ok = B::lmult_inv_jacob_product(Jc, Jp, c, x);
        if (!ok) {
#pragma unroll
          for (int i = 0; i < Y; ++i) x[i] = nan("");
        }
        B::rmult_jacob(Jp, x, delta);
      }
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        if (WANT_MU) mu[k] += delta[k];
        q[k] -= delta[k];
      }
      r.iters += 1;
      r.ndq = update_norm(delta);
    }
    if (WANT_MU) {
#pragma unroll
      for (int k = 0; k < Q; ++k) mu_dt[k] = mu[k] * rdt;
    }
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
    for (int k = 0; k < Q; ++k) mu_dt[k] = (q0[k] - q[k]) * rdt;
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

// want_val = false skips the VALUES of the log-determinant and of the prior density (dim_y / 2
// + 2 fp64 logarithms for the Woodbury family: measured 16 % of a 32-step eight-schools
// launch, profiles/r5b_h_alive.txt); E.nld / E.logdet are then not meaningful.  The
// leapfrog step itself never reads them (systems.py:394-405 uses dh1_dpos only): only the
// Hamiltonian at the ends of a trajectory does.
template <class M>
MLB_DEV bool evaluate_at(const typename M::Params& P, const double (&q)[M::Q], Eval<M>& E,
                         bool want_val = true) {
  double c[M::Y];
  const typename M::Par par = M::params(q[0], q[1]);  // shared by everything below
  M::jacob(P, q, par, E.J, c);
  M::B::decompose_gram(E.J, E.G);
  E.logdet = want_val ? M::B::log_det_sqrt_gram(E.G) : 0.0;
  E.nld = M::nld(P, q, par, E.g, want_val);
  M::add_grad_logdet(P, q, par, E.J, E.G, E.g);
  return E.G.ok;
}

template <int Q>
MLB_DEV double half_sq_norm(const double (&p)[Q]) {
  return 0.5 * dot_acc<Q>(0.0, [&](int k) { return p[k]; }, [&](int k) { return p[k]; });
}

template <class M>
MLB_DEV double hamiltonian(const Eval<M>& E, const double (&p)[M::Q]) {
  return E.nld + E.logdet + half_sq_norm<M::Q>(p);
}

// Up to n_step constrained leapfrog steps A(dt/2) B(dt) A(dt/2) with reversibility check
// (Mici ConstrainedLeapfrogIntegrator) for the chain held by this thread.
//
// Code-size design: one step is the operation sequence
//     [kick, project] solve(+dt) eval  [project] solve(-dt) check  [kick, project]
// i.e. three "sub-phases" that each start with an (optional) half kick and a cotangent
// projection at the current evaluation C, the first two followed by a projection solve
// along the rows of C.J.  The loop below runs sub-phases, so the cotangent projection,
// the projection solver and the evaluation (Jacobian, Gram factor, gradients) are each
// instantiated at ONE code site; the straight-line fully-unrolled register code stays
// inside the instruction cache (the first version inlined 3 + 2 + 2 copies: 64 KB of
// SASS, and ncu showed the warps stalled on instruction fetch).  Only one evaluation is
// live at a time: the forward solve is the last user of the previous point's blocks.
//
// (qw, pw) is the working state in registers.  The last COMMITTED state (start of the
// current step) lives in shared memory at cq / cp (element k at cq[k * kStride]): it is
// what the reversibility check compares against and what a failed step leaves behind.
// PROJECT_FIRST adds a leading sub-phase that projects pw onto the cotangent space at
// the start point (momentum refresh of an HMC transition, systems.py:468-471).
struct TrajOut {
  int status, done, it_fwd, it_rev;
  double h_init, h_last;
};

// run_trajectory_ev: the evaluation C is the caller's.  With have_eval it already holds the
// evaluation at qw (a previous call ended there) and the leading evaluation is skipped;
// after a successful return it holds the evaluation at the new qw.  The dynamic sampler
// chains single steps this way without re-evaluating the Jacobian / Gram factor / gradient
// at every tree leaf.
template <class M, int SOLVER, bool PROJECT_FIRST, bool CHECK_H, int kStride>
MLB_DEV TrajOut run_trajectory_ev(const typename M::Params& P, double (&qw)[M::Q],
                                  double (&pw)[M::Q], double* cq, double* cp, int n_step,
                                  double dt, const Opts& o, double max_delta_h, Eval<M>& C,
                                  bool have_eval, bool want_h = true,
                                  bool project_first = true) {
  using B = typename M::B;
  constexpr int Q = M::Q;
  TrajOut t;
  t.status = MLB_ST_OK, t.done = 0, t.it_fwd = 0, t.it_rev = 0;
  t.h_init = t.h_last = nan("");
  // project_first = false at run time makes a PROJECT_FIRST instantiation behave like the
  // plain one, so that a caller needing both (the dynamic sampler: momentum refresh, then
  // single steps) has ONE copy of the step code instead of two (instruction cache)
  const bool lead = PROJECT_FIRST && project_first;
  int sub = lead ? -1 : 0;
  bool first = true;
  // Interior step boundaries (kFuseKicks): the closing half kick + projection of step s and
  // the opening half kick + projection of step s + 1 act at the same point, and the
  // cotangent projection is linear and idempotent, so P(P(p - dt/2 g) - dt/2 g) =
  // P(p - dt g): one kick and one projection instead of two (what Mici's integrator does
  // step by step differs by rounding only).  The committed momentum is then the one BEFORE
  // the kick (`half_pending`); the failure path finishes the half kick before restoring.
  constexpr bool kFuseKicks = MLB_FUSE_KICKS && !CHECK_H && !PROJECT_FIRST;
  bool half_pending = false;
#pragma unroll
  for (int k = 0; k < Q; ++k) cq[k * kStride] = qw[k], cp[k * kStride] = pw[k];
#pragma unroll 1
  for (;;) {
    if (first || sub == 1) {  // trajectory start / new position after the forward solve
      // energy values: at the trajectory's first point, at every point when the energy is
      // tested per step, otherwise at the last point only (interior commits with fused
      // kicks never form the Hamiltonian)
      const bool want_val =
          want_h && (first || CHECK_H || !kFuseKicks || t.done + 1 >= n_step);
      const bool ok = (first && have_eval) ? C.G.ok : evaluate_at<M>(P, qw, C, want_val);
      if (first && !lead) t.h_init = t.h_last = hamiltonian<M>(C, pw);
      first = false;
      if (!ok) {
        t.status = MLB_ST_LINALG;
        break;
      }
      if (!lead && n_step <= 0) break;
    }
    const bool fuse = kFuseKicks && sub == 2 && t.done + 1 < n_step;
    if (fuse) {
#pragma unroll
      for (int k = 0; k < Q; ++k) cq[k * kStride] = qw[k], cp[k * kStride] = pw[k];
      half_pending = true;
    }
    const double kick = fuse ? -dt : ((sub == 0) || (sub == 2)) ? -0.5 * dt : 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k) pw[k] = kick != 0.0 ? fma(kick, C.g[k], pw[k]) : pw[k];
    B::project_cotangent(C.J, C.G, pw);
    if (PROJECT_FIRST && sub == -1) {
      t.h_init = t.h_last = hamiltonian<M>(C, pw);
#pragma unroll
      for (int k = 0; k < Q; ++k) cp[k * kStride] = pw[k];
      if (n_step <= 0) break;
      sub = 0;
      continue;
    }
    if (fuse) {  // step complete and the next one already opened
      t.done += 1;
      sub = 0;
    } else if (sub == 2) {  // step complete: commit
#pragma unroll
      for (int k = 0; k < Q; ++k) cq[k * kStride] = qw[k], cp[k * kStride] = pw[k];
      half_pending = false;
      t.h_last = hamiltonian<M>(C, pw);
      t.done += 1;
      if (CHECK_H && fabs(t.h_last - t.h_init) > max_delta_h) {
        t.status = MLB_ST_H_DIVERGED;
        break;
      }
      if (t.done >= n_step) break;
      sub = 0;
      continue;
    }
    // B(dt) retraction (sub 0) or the reverse retraction of the reversibility check (sub 1)
    const double sdt = (sub == 0) ? dt : -dt;
    double qs[Q], mu[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) qs[k] = fma(sdt, pw[k], qw[k]);
    // (measured: -4 % launch time, Q fewer live registers and Q fewer DADDs per iteration)
    constexpr bool kMuFromDiff = SOLVER == MLB_SOLVER_QUASI_NEWTON || SOLVER == MLB_SOLVER_NEWTON;
    const SolveOut r = solve_projection<M, SOLVER, !kMuFromDiff>(P, qs, C.J, C.G, sdt, o, mu);
    if (sub == 0) t.it_fwd += r.iters; else t.it_rev += r.iters;
    if (r.status != MLB_ST_OK) {
      t.status = r.status;
      break;
    }
    if (sub == 0) {
      if (kMuFromDiff) {  // sum of the updates = start of the solve - its end
        const double rdt = 1.0 / sdt;
#pragma unroll
        for (int k = 0; k < Q; ++k)
          pw[k] -= (fma(sdt, pw[k], qw[k]) - qs[k]) * rdt, qw[k] = qs[k];
      } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) pw[k] -= mu[k], qw[k] = qs[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < Q; ++k) qs[k] -= cq[k * kStride];
      double rev_err;
      if (MLB_HI_NORM && o.rev_norm == MLB_NORM_MAX) {
        rev_err = max_norm_hi<Q>(qs);
        if (norm_hi_ambiguous(rev_err, o.rev_tol)) rev_err = vec_norm<Q>(qs, MLB_NORM_MAX);
      } else {
        rev_err = vec_norm<Q>(qs, o.rev_norm);
      }
      if (rev_err > o.rev_tol) {
        t.status = MLB_ST_NON_REVERSIBLE;
        break;
      }
    }
    sub += 1;
  }
  if (t.status != MLB_ST_OK && t.status != MLB_ST_H_DIVERGED) {
    // a failed step leaves the chain at its last committed state
#pragma unroll
    for (int k = 0; k < Q; ++k) qw[k] = cq[k * kStride], pw[k] = cp[k * kStride];
    if (kFuseKicks && half_pending) {  // finish the committed step's closing half kick
      evaluate_at<M>(P, qw, C);
#pragma unroll
      for (int k = 0; k < Q; ++k) pw[k] = fma(-0.5 * dt, C.g[k], pw[k]);
      B::project_cotangent(C.J, C.G, pw);
      t.h_last = hamiltonian<M>(C, pw);
    }
  }
  return t;
}

template <class M, int SOLVER, bool PROJECT_FIRST, bool CHECK_H, int kStride>
MLB_DEV TrajOut run_trajectory(const typename M::Params& P, double (&qw)[M::Q],
                               double (&pw)[M::Q], double* cq, double* cp, int n_step,
                               double dt, const Opts& o, double max_delta_h,
                               bool want_h = true) {
  Eval<M> C;
  return run_trajectory_ev<M, SOLVER, PROJECT_FIRST, CHECK_H, kStride>(
      P, qw, pw, cq, cp, n_step, dt, o, max_delta_h, C, false, want_h);
}

}  // namespace mlb
