// Two warps per batch of 32 chains for SMALL launches of the register-resident family.
//
// Below ~16,384 chains a B200 holds at most one warp per scheduler and a launch of the
// fused trajectory kernel takes what ONE chain's instruction stream takes (0.25 ms for 32
// eight-schools steps whatever the batch: DESIGN 5) - the regime of the 8-GPU split of the
// BASELINE configuration (8,192 chains per GPU) and of the chunks of the host entry point.
// A step is  [kick, project] solve(+dt) eval [project]  solve(-dt) check  [kick, project],
// and the reverse solve of the reversibility check (Mici ConstrainedLeapfrogIntegrator,
// SURVEY 3.2) only produces a verdict: nothing of step s + 1 needs its result.  So it runs
// on a SECOND warp, one step behind the first:
//
//   driver warp    front(s): opening kick + projection, forward solve, evaluation at the new
//                  point, projection of the momentum there; publishes (q_new, p, q_old)
//   checker warp   check(s): Jacobian (and Gram factor) at q_new - the same device functions
//                  the evaluation uses, so the same bits -, reverse solve, |q_back - q_old|
//   driver warp    back(s): closing kick -> candidate state of step s; speculatively goes
//                  on to front(s + 1) while check(s) runs
//
// with ONE __syncthreads per step (exchange block and verdicts double-buffered).  A step is
// committed when its verdict arrives (one step later); a failed verdict discards the
// speculative work and its iteration counts, so states, statuses, completed steps and
// iteration counts are those of the sequential integrator.  The chain's critical path per
// step drops from (forward solve + evaluation + kicks + reverse solve) to the first three.
// Interior half kicks are fused as in the one-warp kernels (one kick of dt and one projection
// at a step boundary); the candidate state of step s, which has to exist before its verdict,
// takes its momentum as the mean of the momenta before and after that kick.
#pragma once
// (included by mlb_launch.cuh after load_vec / store_vec)

namespace mlb {

template <class M>
constexpr int duo_smem_bytes() {
  return (int)(sizeof(double) * 10 * M::Q * 32 + sizeof(int32_t) * 6 * 32);
}

template <class M, int SOLVER>
__global__ void __launch_bounds__(64, 1)
    k_leapfrog_steps_duo(const __grid_constant__ typename M::Params P, int64_t n, int64_t es,
                         int64_t cs, double* q, double* p, double dt, const double* dt_pc,
                         int n_step, Opts o, int32_t* status, int32_t* n_done,
                         int32_t* iters_fwd, int32_t* iters_rev, double* h_init,
                         double* h_final) {
  using B = typename M::B;
  constexpr int Q = M::Q, Y = M::Y;
  constexpr bool kMuFromDiff = SOLVER == MLB_SOLVER_QUASI_NEWTON || SOLVER == MLB_SOLVER_NEWTON;
  constexpr bool kFuse = MLB_FUSE_KICKS != 0;  // fused interior half kicks (see back(s))
  extern __shared__ double duo_sm[];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;  // 0 driver, 1 checker
  double* const col = duo_sm + lane;                            // element r at col[r * 32]
  // rows: exchange block s & 1 at 3 Q (s & 1): q_new, p (projected at q_new), q_old;
  // verified state VQ VP; candidate state CQ CP
  constexpr int XQN = 0, XPP = Q, XQP = 2 * Q, XSZ = 3 * Q, VQ = 6 * Q, VP = 7 * Q, CQ = 8 * Q,
                CP = 9 * Q;
  int32_t* const isl = reinterpret_cast<int32_t*>(duo_sm + 10 * Q * 32) + lane;
  // int rows: published[2], verdict[2], verdict iterations[2]
  constexpr int IPUB = 0, IVER = 2, IVIT = 4;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = i < n;
  const double dti = valid ? (dt_pc ? dt_pc[i] : dt) : 0.0;
  const bool want_h = h_init != nullptr || h_final != nullptr;

  // ---- driver state
  double qw[Q], pw[Q];
  Eval<M> C;
  bool alive = false;
  int st = MLB_ST_OK, done = 0, it_f = 0, it_r = 0, spec_it = 0, front_st = MLB_ST_OK;
  double h0 = nan(""), h_ver = nan(""), h_cand = nan("");
  if (role == 0 && valid) {
    load_vec<Q>(qw, q, es, cs, i);
    load_vec<Q>(pw, p, es, cs, i);
    const bool ok = evaluate_at<M>(P, qw, C, want_h);
    h0 = h_ver = hamiltonian<M>(C, pw);
#pragma unroll
    for (int k = 0; k < Q; ++k) col[(VQ + k) * 32] = qw[k], col[(VP + k) * 32] = pw[k];
    alive = ok && n_step > 0;
    if (!ok) st = MLB_ST_LINALG;
  }
#pragma unroll 1
  for (int s = 0; s <= n_step; ++s) {
    double* const X = col + (s & 1) * XSZ * 32;
    if (role == 0) {
      if (s < n_step) {
        int pub = 0;
        if (alive) {  // front(s): everything up to the momentum projected at the new point
          if (!kFuse || s == 0) {  // (else the opening kick was fused into back(s - 1))
#pragma unroll
            for (int k = 0; k < Q; ++k) pw[k] = fma(-0.5 * dti, C.g[k], pw[k]);
            B::project_cotangent(C.J, C.G, pw);
          }
          double qs[Q], mu[Q];
#pragma unroll
          for (int k = 0; k < Q; ++k) qs[k] = fma(dti, pw[k], qw[k]);
          const SolveOut r =
              solve_projection<M, SOLVER, !kMuFromDiff>(P, qs, C.J, C.G, dti, o, mu);
          spec_it = r.iters, front_st = r.status;
          if (r.status == MLB_ST_OK) {
#pragma unroll
            for (int k = 0; k < Q; ++k) X[(XQP + k) * 32] = qw[k];
            if (kMuFromDiff) {
              const double rdt = 1.0 / dti;
#pragma unroll
              for (int k = 0; k < Q; ++k)
                pw[k] -= (fma(dti, pw[k], qw[k]) - qs[k]) * rdt, qw[k] = qs[k];
            } else {
#pragma unroll
              for (int k = 0; k < Q; ++k) pw[k] -= mu[k], qw[k] = qs[k];
            }
            if (!evaluate_at<M>(P, qw, C, want_h)) {
              front_st = MLB_ST_LINALG;
            } else {
              B::project_cotangent(C.J, C.G, pw);
#pragma unroll
              for (int k = 0; k < Q; ++k) X[(XQN + k) * 32] = qw[k], X[(XPP + k) * 32] = pw[k];
              pub = 1;
            }
          }
        }
        isl[(IPUB + (s & 1)) * 32] = pub;
      }
    }
    __syncthreads();
    if (role == 0) {
      if (alive && s >= 1) {  // verdict of step s - 1
        const int v = isl[(IVER + ((s - 1) & 1)) * 32];
        it_r += isl[(IVIT + ((s - 1) & 1)) * 32];
        if (v != MLB_ST_OK) {
          st = v, alive = false;  // the chain stays at its verified state
        } else {
#pragma unroll
          for (int k = 0; k < Q; ++k)
            col[(VQ + k) * 32] = col[(CQ + k) * 32], col[(VP + k) * 32] = col[(CP + k) * 32];
          h_ver = h_cand, done = s;
        }
      }
      if (alive && s < n_step) {  // back(s)
        it_f += spec_it;
        if (front_st != MLB_ST_OK) {
          st = front_st, alive = false;
        } else {
          if (kFuse && s + 1 < n_step) {
            // interior boundary: ONE kick of dt and one projection open step s + 1; the
            // candidate momentum of step s, P(p - dt/2 g) = p - dt/2 P g (the projection is
            // linear and p is already tangent), is the mean of p and that
            double pc[Q];
#pragma unroll
            for (int k = 0; k < Q; ++k) pc[k] = pw[k], pw[k] = fma(-dti, C.g[k], pw[k]);
            B::project_cotangent(C.J, C.G, pw);
#pragma unroll
            for (int k = 0; k < Q; ++k) {
              pc[k] = 0.5 * (pc[k] + pw[k]);
              col[(CQ + k) * 32] = qw[k], col[(CP + k) * 32] = pc[k];
            }
            h_cand = want_h ? hamiltonian<M>(C, pc) : 0.0;
          } else {
#pragma unroll
            for (int k = 0; k < Q; ++k) pw[k] = fma(-0.5 * dti, C.g[k], pw[k]);
            B::project_cotangent(C.J, C.G, pw);
#pragma unroll
            for (int k = 0; k < Q; ++k) col[(CQ + k) * 32] = qw[k], col[(CP + k) * 32] = pw[k];
            h_cand = hamiltonian<M>(C, pw);
          }
        }
      }
    } else if (s < n_step && valid && isl[(IPUB + (s & 1)) * 32]) {
      // check(s): reverse retraction from the new point along its own Jacobian rows
      double qn[Q], qb[Q], mu[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) qn[k] = X[(XQN + k) * 32];
      typename B::Jac Jn;
      typename B::Gram Gn;
      {
        double c[Y];
        M::jacob(P, qn, Jn, c);
      }
      if (SOLVER == MLB_SOLVER_QUASI_NEWTON || SOLVER == MLB_SOLVER_MICI_QUASI_NEWTON)
        B::decompose_gram(Jn, Gn);
#pragma unroll
      for (int k = 0; k < Q; ++k) qb[k] = fma(-dti, X[(XPP + k) * 32], qn[k]);
      const SolveOut r = solve_projection<M, SOLVER, !kMuFromDiff>(P, qb, Jn, Gn, -dti, o, mu);
      int v = r.status;
      if (v == MLB_ST_OK) {
#pragma unroll
        for (int k = 0; k < Q; ++k) qb[k] -= X[(XQP + k) * 32];
        double rev_err;
        if (MLB_HI_NORM && o.rev_norm == MLB_NORM_MAX) {
          rev_err = max_norm_hi<Q>(qb);
          if (norm_hi_ambiguous(rev_err, o.rev_tol)) rev_err = vec_norm<Q>(qb, MLB_NORM_MAX);
        } else {
          rev_err = vec_norm<Q>(qb, o.rev_norm);
        }
        if (rev_err > o.rev_tol) v = MLB_ST_NON_REVERSIBLE;
      }
      isl[(IVER + (s & 1)) * 32] = v, isl[(IVIT + (s & 1)) * 32] = r.iters;
    }
  }
  if (role == 0 && valid) {
    double qo[Q], po[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) qo[k] = col[(VQ + k) * 32], po[k] = col[(VP + k) * 32];
    store_vec<Q>(qo, q, es, cs, i);
    store_vec<Q>(po, p, es, cs, i);
    status[i] = st;
    if (n_done) n_done[i] = done;
    if (iters_fwd) iters_fwd[i] = it_f;
    if (iters_rev) iters_rev[i] = it_r;
    if (h_init) h_init[i] = h0;
    if (h_final) h_final[i] = h_ver;
  }
}

}  // namespace mlb
