// Occupancy-oriented form of the fused trajectory for the hierarchical (eight-schools-shaped)
// model: the SAME device functions as mlb_chmc.cuh (evaluate_at, Blocks::project_cotangent,
// solve_projection - hence the same arithmetic), but the chain's persistent state lives in a
// per-thread column of SHARED memory and each sub-phase loads only what it needs into
// registers and stores its results back.
//
// Why: the register-resident kernel (k_leapfrog_steps) needs 255 registers (plus spills) for
// the union of everything that is live across a step - working (q, p), the evaluation at
// the current point (Jacobian blocks, Gram factor, gradient), the trial position of the
// solver - which limits an SM to 4 CTAs x 2 warps: two warps per scheduler, each issuing a
// dependent fp64 chain (DFMA latency 8.3 cycles, issue interval 2.1: profiles/
// r4b_ubench_fp64.txt), and 65,536 chains need 1.73 waves.  No single sub-phase needs more
// than ~128 registers, so with the state parked between sub-phases the kernel is capped at
// 144 registers, 7 CTAs (14 warps) per SM, and 65,536 chains are ONE wave of 1,024 CTAs on
// 148 x 7 = 1,036 slots.  Per-thread shared column: 14 + 6 dim_y doubles (62 for dim_y = 8,
// 31 KB per 64-thread CTA, 7 CTAs = 217 KB of the SM's 227 KB).  The last committed state of
// a chain is the global (q, p) arrays themselves (written at every completed step, read back
// for the reversibility check), not a second on-chip copy.
#pragma once
#include "mlb_chmc.cuh"

namespace mlb {

template <class M, int KB>
struct LeanCol {
  static constexpr int Y = M::Y, Q = M::Q;
  // slots of the column (element k at s[k * KB])
  static constexpr int SQ = 0, SP = Q, SG = 2 * Q, SE = SG + 2 + Y, SRD = SE + 7,
                       ROWS = SRD + Y;
  enum { E_TAU, E_DTAU, E_L00, E_L10, E_L11, E_R0, E_R1 };
  double* s;
  MLB_DEV double& operator[](int k) const { return s[k * KB]; }
};

// phases must not forward register copies of the parked state to one another
#define MLB_LEAN_FENCE() asm volatile("" ::: "memory")

template <class M, int SOLVER, int KB>
struct HierLean {
  using B = typename M::B;
  using C = LeanCol<M, KB>;
  static constexpr int Y = M::Y, Q = M::Q;
  static_assert(B::FAM == FAM_HIER && !B::DENSE && B::U == 2, "hierarchical Woodbury family");

  // Jacobian blocks / Gram factor at the parked point, rebuilt from the column
  static MLB_DEV void load_jac(const typename M::Params& P, C S, typename B::Jac& J) {
    const double tau = S[C::SE + C::E_TAU], dtau = S[C::SE + C::E_DTAU];
    const double dmu = M::params(0.0, 0.0).dmu;  // a constant of the parametrisation
#pragma unroll
    for (int i = 0; i < Y; ++i) {
      J.du[i][0] = dmu, J.du[i][1] = dtau * S[C::SQ + 2 + i];
      J.dv[i] = tau, J.dn[i] = P.sigma[i];
    }
  }
  static MLB_DEV void load_gram(C S, typename B::Gram& G) {
    G.chol[0][0] = S[C::SE + C::E_L00], G.chol[0][1] = 0.0;
    G.chol[1][0] = S[C::SE + C::E_L10], G.chol[1][1] = S[C::SE + C::E_L11];
    G.rdiag[0] = S[C::SE + C::E_R0], G.rdiag[1] = S[C::SE + C::E_R1];
#pragma unroll
    for (int i = 0; i < Y; ++i) G.rd[i] = S[C::SRD + i], G.d[i] = 0.0;  // d: log det only
    G.ok = true;
  }

  // evaluation at the parked position: parks tau, dtau, Gram factor, 1 / d and the part of
  // dh1/dq that is not the position itself (d nld / d n_i = n_i); returns nld + log det
  static MLB_DEV bool evaluate(const typename M::Params& P, C S, double& h1,
                               bool want_val = true) {
    double q[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) q[k] = S[C::SQ + k];
    Eval<M> E;
    const bool ok = evaluate_at<M>(P, q, E, want_val);
    h1 = E.nld + E.logdet;
#pragma unroll
    for (int k = 0; k < 2 + Y; ++k) S[C::SG + k] = E.g[k];
    S[C::SE + C::E_TAU] = E.J.dv[0];
    S[C::SE + C::E_DTAU] = M::params(q[0], q[1]).dtau;
    S[C::SE + C::E_L00] = E.G.chol[0][0], S[C::SE + C::E_L10] = E.G.chol[1][0];
    S[C::SE + C::E_L11] = E.G.chol[1][1];
    S[C::SE + C::E_R0] = E.G.rdiag[0], S[C::SE + C::E_R1] = E.G.rdiag[1];
#pragma unroll
    for (int i = 0; i < Y; ++i) S[C::SRD + i] = E.G.rd[i];
    return ok;
  }

  // p <- P_q (p + kick * dh1/dq) at the parked evaluation; returns |p|^2 / 2
  static MLB_DEV double kick_project(const typename M::Params& P, C S, double kick) {
    double p[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) p[k] = S[C::SP + k];
    if (kick != 0.0) {
#pragma unroll
      for (int k = 0; k < 2 + Y; ++k) p[k] = fma(kick, S[C::SG + k], p[k]);
#pragma unroll
      for (int i = 0; i < Y; ++i) p[2 + Y + i] = fma(kick, S[C::SQ + 2 + Y + i], p[2 + Y + i]);
    }
    typename B::Jac J;
    typename B::Gram G;
    load_jac(P, S, J);
    load_gram(S, G);
    B::project_cotangent(J, G, p);
#pragma unroll
    for (int k = 0; k < Q; ++k) S[C::SP + k] = p[k];
    return half_sq_norm<Q>(p);
  }
};

// n_step constrained leapfrog steps for the chain whose column is S and whose committed
// state is (gq, gp) in global memory (element k at gq[k * es]); semantics of
// run_trajectory<M, SOLVER, false, false> with fused interior kicks (mlb_chmc.cuh).
template <class M, int SOLVER, int KB>
MLB_DEV TrajOut run_trajectory_lean(const typename M::Params& P, LeanCol<M, KB> S, double* gq,
                                    double* gp, int64_t es, int n_step, double dt,
                                    const Opts& o, bool want_h = true,
                                    bool preloaded = false) {
  using H = HierLean<M, SOLVER, KB>;
  using B = typename M::B;
  using C = LeanCol<M, KB>;
  constexpr int Q = M::Q;
  TrajOut t;
  t.status = MLB_ST_OK, t.done = 0, t.it_fwd = 0, t.it_rev = 0;
  t.h_init = t.h_last = nan("");
  if (!preloaded) {  // else the kernel staged (q, p) into the column by bulk copies
#pragma unroll
    for (int k = 0; k < Q; ++k) S[C::SQ + k] = gq[k * es], S[C::SP + k] = gp[k * es];
  }
  double h1;
  bool half_pending = false;  // the committed momentum still lacks its closing half kick
  {
    const bool ok = H::evaluate(P, S, h1, want_h);
    double ke = 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k) ke = fma(S[C::SP + k], S[C::SP + k], ke);
    t.h_init = t.h_last = h1 + 0.5 * ke;
    if (!ok) t.status = MLB_ST_LINALG;
  }
  MLB_LEAN_FENCE();
  int sub = 0;
  while (t.status == MLB_ST_OK && t.done < n_step) {
    if (sub == 1) {  // new position after the forward solve
      // energy values at the last point only (see evaluate_at)
      if (!H::evaluate(P, S, h1, want_h && (!MLB_FUSE_KICKS || t.done + 1 >= n_step))) {
        t.status = MLB_ST_LINALG;
        break;
      }
      MLB_LEAN_FENCE();
    }
    const bool fuse = MLB_FUSE_KICKS && sub == 2 && t.done + 1 < n_step;
    if (fuse) {  // commit the step with its momentum BEFORE the closing kick
#pragma unroll
      for (int k = 0; k < Q; ++k) gq[k * es] = S[C::SQ + k], gp[k * es] = S[C::SP + k];
      half_pending = true;
    }
    const double kick = fuse ? -dt : (sub == 1 ? 0.0 : -0.5 * dt);
    const double ke = H::kick_project(P, S, kick);
    MLB_LEAN_FENCE();
    if (fuse) {
      t.done += 1;
      sub = 0;
    } else if (sub == 2) {
#pragma unroll
      for (int k = 0; k < Q; ++k) gq[k * es] = S[C::SQ + k], gp[k * es] = S[C::SP + k];
      half_pending = false;
      t.h_last = h1 + ke;
      t.done += 1;
      sub = 0;
      continue;
    }
    // B(dt) retraction (sub 0) or the reverse retraction of the reversibility check (sub 1)
    const double sdt = sub == 0 ? dt : -dt;
    double qs[Q], mu[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) qs[k] = fma(sdt, S[C::SP + k], S[C::SQ + k]);
    SolveOut r;
    {
      typename B::Jac Jp;
      typename B::Gram Gp;
      H::load_jac(P, S, Jp);
      if (SOLVER == MLB_SOLVER_QUASI_NEWTON) H::load_gram(S, Gp);
      constexpr bool kMuFromDiff =
          SOLVER == MLB_SOLVER_QUASI_NEWTON || SOLVER == MLB_SOLVER_NEWTON;
      r = solve_projection<M, SOLVER, !kMuFromDiff>(P, qs, Jp, Gp, sdt, o, mu);
      if (sub == 0) t.it_fwd += r.iters; else t.it_rev += r.iters;
      if (r.status != MLB_ST_OK) {
        t.status = r.status;
        break;
      }
      if (sub == 0) {
        const double rdt = 1.0 / sdt;
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const double pk = S[C::SP + k];
          const double m = kMuFromDiff ? (fma(sdt, pk, S[C::SQ + k]) - qs[k]) * rdt : mu[k];
          S[C::SP + k] = pk - m, S[C::SQ + k] = qs[k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) qs[k] -= gq[k * es];
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
    }
    MLB_LEAN_FENCE();
    sub += 1;
  }
  if (t.status != MLB_ST_OK && half_pending) {
    // the chain stays at its last committed step: finish that step's closing half kick
#pragma unroll
    for (int k = 0; k < Q; ++k) S[C::SQ + k] = gq[k * es], S[C::SP + k] = gp[k * es];
    H::evaluate(P, S, h1, want_h);
    MLB_LEAN_FENCE();
    const double ke = H::kick_project(P, S, -0.5 * dt);
    t.h_last = h1 + ke;
#pragma unroll
    for (int k = 0; k < Q; ++k) gp[k * es] = S[C::SP + k];
  }
  return t;
}

}  // namespace mlb
