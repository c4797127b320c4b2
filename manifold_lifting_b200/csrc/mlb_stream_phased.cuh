// Phase-structured constrained leapfrog for the streaming (ODE-constrained) family.
//
// The monolithic per-chain kernel (mlb_stream_launch.cuh) leaves a B200 almost empty: a
// FitzHugh-Nagumo batch of 16,384 chains is 512 warps on 592 sub-partitions, and each warp
// walks its sweeps one sensitivity direction / pair group after another.  All chains of a
// launch go through the SAME sequence of sub-phases (only solver iteration counts differ
// per chain), and every Y-sized intermediate already lives in global scratch, so the
// trajectory is re-cut into kernels at sub-phase boundaries and driven from the host:
//
//   * the expensive sweeps run on a 2-D grid (chain block x ROLE): a Jacobian sweep is
//     split by sensitivity direction (FN 4 roles x 2 directions, HH 6 x 1), the
//     second-order sweep by pair group (FN 6, HH 7) - 4-7x more resident warps, no
//     redundant arithmetic beyond the state itself, small per-kernel code;
//   * per-chain control (iteration counters, norms, status, alive / active masks) sits in
//     small device arrays; the solver loop is a host loop that reads ONE int (chains still
//     iterating) per iteration - microseconds against millisecond sweeps; warps whose 32
//     chains have all converged exit at once;
//   * arithmetic is that of the monolithic kernel (same device functions), so the two
//     paths agree bitwise and share the parity tests.
#pragma once
#include "mlb_stream_launch.cuh"

namespace mlb {

// per-chain control block rows in the scratch allocation (after StreamScratch::rows)
enum { CT_ALIVE, CT_ACTIVE, CT_ITERS, CT_STATUS, CT_DONE, CT_TF, CT_TR, CT_NCONSTR, CT_INT_ROWS };
enum { CD_NDQ, CD_ERR, CD_NLD, CD_LOGDET, CD_HINIT, CD_HLAST, CD_DBL_ROWS };

struct Phased {
  StreamScratch S;
  Col hpart;     // [NG * ND] partial second-order gradients
  double* cd;    // [CD_DBL_ROWS][lds]
  int32_t* ct;   // [CT_INT_ROWS][lds]
  int64_t lds;
  MLB_DEV int32_t& I(int row, int64_t i) const { return ct[(int64_t)row * lds + i]; }
  MLB_DEV double& Dv(int row, int64_t i) const { return cd[(int64_t)row * lds + i]; }
};

struct PhasedLayout {
  double* base;
  int64_t lds;
  int U, Y, NGND;
  int64_t rows_main() const { return StreamScratch::rows(U, Y); }
  int64_t rows_total() const {
    return rows_main() + NGND + CD_DBL_ROWS + (CT_INT_ROWS + 1) / 2 + 1;
  }
  __host__ __device__ Phased at(int64_t i) const {
    Phased p;
    p.lds = lds;
    double* after = base + StreamScratch::rows(U, Y) * lds;
    p.hpart = Col{after + i, lds};
    p.cd = after + (int64_t)NGND * lds;
    p.ct = reinterpret_cast<int32_t*>(p.cd + (int64_t)CD_DBL_ROWS * lds);
#ifdef __CUDA_ARCH__
    p.S = StreamScratch::make(base, lds, i, U, Y);
#endif
    return p;
  }
};

#define MLB_PH_INDEX()                                              \
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
  if (i >= n) return;                                               \
  const Phased P = L.at(i)

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_begin(PhasedLayout L, int64_t n, int64_t ld, const double* q, const double* p) {
  MLB_PH_INDEX();
  const int Q = L.U + L.Y;
  const Col qw = col(q, ld, i), pw = col(p, ld, i);
#pragma unroll 4
  for (int k = 0; k < Q; ++k) P.S.cq[k] = qw[k], P.S.cp[k] = pw[k];
  P.I(CT_ALIVE, i) = 1, P.I(CT_ACTIVE, i) = 0, P.I(CT_ITERS, i) = 0;
  P.I(CT_STATUS, i) = MLB_ST_OK, P.I(CT_DONE, i) = 0, P.I(CT_TF, i) = 0, P.I(CT_TR, i) = 0;
  P.I(CT_NCONSTR, i) = 0;
  P.Dv(CD_HINIT, i) = nan(""), P.Dv(CD_HLAST, i) = nan("");
}

// Jacobian sweep, one role per blockIdx.y.  src / dst select the position (working qw or
// trial qs) and the destination blocks (Jc at an evaluation, Jt inside Newton).
template <class M, bool EVAL>
__global__ void __launch_bounds__(kStreamBlock)
    kp_jacob(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n, int64_t ld,
             const double* q) {
  MLB_PH_INDEX();
  if (!P.I(EVAL ? CT_ALIVE : CT_ACTIVE, i)) return;
  const Col src = EVAL ? col(q, ld, i) : P.S.qs;
  const Col dst = EVAL ? P.S.Jc : P.S.Jt;
  const Col c = EVAL ? Col{nullptr, 0} : P.S.c;
  constexpr int R = M::ROLE_DIRS;
  [&]<int... K>(std::integer_sequence<int, K...>) {
    ((blockIdx.y == K ? Stream<M>::template jacob_role<K * R, R, K == 0>(D, src, dst, c)
                      : (void)0),
     ...);
  }(std::make_integer_sequence<int, M::N_ROLE>{});
}

// residual sweep at the trial position (quasi-Newton)
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_constr(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n) {
  MLB_PH_INDEX();
  if (!P.I(CT_ACTIVE, i)) return;
  double u[M::U];
  Stream<M>::load_u(P.S.qs, u);
  const Col qs = P.S.qs;
  Stream<M>::constr(D, u, [&](int k) { return qs[M::U + k]; }, P.S.c, MLB_NORM_MAX);
}

// after the Jacobian sweep of an evaluation: Gram factor, log det, prior, and everything
// of grad 0.5 log det(J J^T) that does not need second-order sensitivities
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_eval_mid(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n, int64_t ld,
                double* q, double* p) {
  MLB_PH_INDEX();
  if (!P.I(CT_ALIVE, i)) return;
  using S = Stream<M>;
  constexpr int U = M::U;
  const Col qw = col(q, ld, i);
  typename S::Chol C;
  const bool ok = S::decompose(D.Y, P.S.Jc, P.S.Gc, C);
  const double ldv = S::logdet(D.Y, P.S.Gc, C);
  double u[U], gu[U];
  S::load_u(qw, u);
  double nld = M::nld_u(u, gu);
  S::template grad_logdet<false>(D, qw, u, P.S.Jc, C, P.S.Jt, gu, P.S.g, nld);
#pragma unroll
  for (int a = 0; a < U; ++a) P.S.g[a] = gu[a];
  P.Dv(CD_NLD, i) = nld, P.Dv(CD_LOGDET, i) = ldv;
  if (!ok) {  // hand the chain back at its last committed state
    P.I(CT_STATUS, i) = MLB_ST_LINALG, P.I(CT_ALIVE, i) = 0;
    const Col pw = col(p, ld, i);
    const int Q = U + D.Y;
#pragma unroll 4
    for (int k = 0; k < Q; ++k) qw[k] = P.S.cq[k], pw[k] = P.S.cp[k];
  }
}

// second-order sweep, one pair group per blockIdx.y; partial gradients to hpart
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_second_order(const __grid_constant__ OdeData D, PhasedLayout L, int64_t n, int64_t ld,
                    const double* q) {
  MLB_PH_INDEX();
  if (!P.I(CT_ALIVE, i)) return;
  constexpr int U = M::U;
  double u[U], h[U];
  Stream<M>::load_u(col(q, ld, i), u);
#pragma unroll
  for (int a = 0; a < U; ++a) h[a] = 0.0;
  [&]<int... G>(std::integer_sequence<int, G...>) {
    ((blockIdx.y == G ? Stream<M>::template second_order_group<G>(D, u, P.S.Jt, h) : (void)0),
     ...);
  }(std::make_integer_sequence<int, M::NG>{});
#pragma unroll
  for (int d = 0; d < M::ND; ++d) P.hpart[(int)blockIdx.y * M::ND + d] = h[M::dir_u(d)];
}

// [finish a fresh evaluation: add the second-order partials] [half kick] cotangent
// projection [commit the step / record h_init]
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_kick_project(PhasedLayout L, int64_t n, int64_t ld, double* p, double* q, double dt,
                    const double* dt_pc, int fresh, int kick, int first, int commit) {
  MLB_PH_INDEX();
  if (!P.I(CT_ALIVE, i)) return;
  using S = Stream<M>;
  constexpr int U = M::U;
  const int Y = L.Y, Q = U + Y;
  const Col pw = col(p, ld, i), qw = col(q, ld, i);
  if (fresh) {
#pragma unroll
    for (int d = 0; d < M::ND; ++d) {
      double s = P.S.g[M::dir_u(d)];
      for (int g = 0; g < M::NG; ++g) s += P.hpart[g * M::ND + d];
      P.S.g[M::dir_u(d)] = s;
    }
  }
  auto hamiltonian = [&]() {
    double s = 0.0;
#pragma unroll 4
    for (int k = 0; k < Q; ++k) s = fma(pw[k], pw[k], s);
    return P.Dv(CD_NLD, i) + P.Dv(CD_LOGDET, i) + 0.5 * s;
  };
  if (first) P.Dv(CD_HINIT, i) = P.Dv(CD_HLAST, i) = hamiltonian();
  const double dti = dt_pc ? dt_pc[i] : dt;
  if (kick) {
#pragma unroll 4
    for (int k = 0; k < Q; ++k) pw[k] = fma(-0.5 * dti, P.S.g[k], pw[k]);
  }
  typename S::Chol C;
  S::load_chol(P.S.Gc, C);
  S::project(Y, P.S.Jc, C, pw, P.S.w);
  if (commit) {
#pragma unroll 4
    for (int k = 0; k < Q; ++k) P.S.cq[k] = qw[k], P.S.cp[k] = pw[k];
    P.Dv(CD_HLAST, i) = hamiltonian();
    P.I(CT_DONE, i) += 1;
  }
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_solve_begin(PhasedLayout L, int64_t n, int64_t ld, const double* q, const double* p,
                   double dt, const double* dt_pc, double sign, int max_iters, int* counter) {
  MLB_PH_INDEX();
  const int alive = P.I(CT_ALIVE, i);
  const bool go = alive && max_iters > 0;
  P.I(CT_ACTIVE, i) = go, P.I(CT_ITERS, i) = 0, P.I(CT_NCONSTR, i) = 0;
  P.Dv(CD_NDQ, i) = INFINITY, P.Dv(CD_ERR, i) = -1.0;
  if (!alive) return;
  const int Q = L.U + L.Y;
  const double sdt = sign * (dt_pc ? dt_pc[i] : dt);
  const Col qw = col(q, ld, i), pw = col(p, ld, i);
#pragma unroll 4
  for (int k = 0; k < Q; ++k) P.S.qs[k] = fma(sdt, pw[k], qw[k]), P.S.mu[k] = 0.0;
  if (go) atomicAdd(counter, 1);
}

// one solver iteration after its sweep: |c|, direction, update, new |dq|, and whether the
// chain iterates again (the loop condition of systems.py:215-227 for the NEXT pass)
template <class M, int SOLVER>
__global__ void __launch_bounds__(kStreamBlock)
    kp_update(PhasedLayout L, int64_t n, Opts o, int* counter) {
  MLB_PH_INDEX();
  if (!P.I(CT_ACTIVE, i)) return;
  using S = Stream<M>;
  constexpr int U = M::U;
  const int Y = L.Y;
  const Col q = P.S.qs, mu = P.S.mu, c = P.S.c, jp = P.S.Jc;
  const Col jl = SOLVER == MLB_SOLVER_QUASI_NEWTON ? P.S.Jc : P.S.Jt;
  double s[U], du[U], u[U];
  S::load_u(q, u);
  double eacc = 0.0;
#pragma unroll 4
  for (int k = 0; k < Y; ++k) eacc = norm_accum(eacc, c[k], o.norm);
  const double err = norm_finish(eacc, o.norm);
  if (SOLVER == MLB_SOLVER_QUASI_NEWTON) {
    typename S::Chol Cp;
    S::load_chol(P.S.Gc, Cp);
#pragma unroll
    for (int a = 0; a < U; ++a) s[a] = 0.0;
#pragma unroll 4
    for (int k = 0; k < Y; ++k) {
      const double dn = jp[U * Y + k], cr = c[k] * rcp_fast(dn * dn);
#pragma unroll
      for (int a = 0; a < U; ++a) s[a] = fma(jp[k * U + a], cr, s[a]);
    }
    chol_solve<U>(Cp.L, Cp.rdiag, s);
  } else {
    if (!S::inv_jacprod_core(Y, jl, jp, c, s)) {
#pragma unroll
      for (int a = 0; a < U; ++a) s[a] = nan("");
    }
  }
#pragma unroll
  for (int a = 0; a < U; ++a) du[a] = 0.0;
  double nacc = 0.0;
#pragma unroll 4
  for (int k = 0; k < Y; ++k) {
    const double x = S::inv_jacprod_x(Y, k, jl, jp, c[k], s);
#pragma unroll
    for (int a = 0; a < U; ++a) du[a] = fma(jp[k * U + a], x, du[a]);
    const double dni = jp[U * Y + k] * x;
    q[U + k] -= dni, mu[U + k] += dni;
    nacc = norm_accum(nacc, dni, o.norm);
  }
#pragma unroll
  for (int a = 0; a < U; ++a) {
    q[a] = u[a] - du[a];
    mu[a] += du[a];
    nacc = norm_accum(nacc, du[a], o.norm);
  }
  const double ndq = norm_finish(nacc, o.norm);
  const int iters = P.I(CT_ITERS, i) + 1;
  P.I(CT_ITERS, i) = iters, P.Dv(CD_NDQ, i) = ndq, P.Dv(CD_ERR, i) = err;
  const bool again = keep_iterating(iters, ndq, err, o);
  P.I(CT_ACTIVE, i) = again;
  if (again) atomicAdd(counter, 1);
}

// close a solve: mu / dt, classification (solvers.py:81-95), then either accept the
// retraction (forward) or run the reversibility check (reverse); failures restore the
// last committed state
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_solve_end(PhasedLayout L, int64_t n, int64_t ld, double* q, double* p, double dt,
                 const double* dt_pc, double sign, Opts o, int forward) {
  MLB_PH_INDEX();
  if (!P.I(CT_ALIVE, i)) return;
  const int Q = L.U + L.Y;
  const Col qw = col(q, ld, i), pw = col(p, ld, i);
  const double rdt = 1.0 / (sign * (dt_pc ? dt_pc[i] : dt));
  P.I(forward ? CT_TF : CT_TR, i) += P.I(CT_ITERS, i);
  int st = classify(P.Dv(CD_ERR, i), P.Dv(CD_NDQ, i), o);
  if (st == MLB_ST_OK) {
    if (forward) {
#pragma unroll 4
      for (int k = 0; k < Q; ++k) pw[k] = fma(-rdt, P.S.mu[k], pw[k]), qw[k] = P.S.qs[k];
    } else {
      double acc = 0.0;
#pragma unroll 4
      for (int k = 0; k < Q; ++k) acc = norm_accum(acc, P.S.qs[k] - P.S.cq[k], o.rev_norm);
      if (norm_finish(acc, o.rev_norm) > o.rev_tol) st = MLB_ST_NON_REVERSIBLE;
    }
  }
  if (st != MLB_ST_OK) {
    P.I(CT_STATUS, i) = st, P.I(CT_ALIVE, i) = 0;
#pragma unroll 4
    for (int k = 0; k < Q; ++k) qw[k] = P.S.cq[k], pw[k] = P.S.cp[k];
  }
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    kp_finish(PhasedLayout L, int64_t n, int32_t* status, int32_t* n_done, int32_t* iters_fwd,
              int32_t* iters_rev, double* h_init, double* h_final) {
  MLB_PH_INDEX();
  status[i] = P.I(CT_STATUS, i);
  if (n_done) n_done[i] = P.I(CT_DONE, i);
  if (iters_fwd) iters_fwd[i] = P.I(CT_TF, i);
  if (iters_rev) iters_rev[i] = P.I(CT_TR, i);
  if (h_init) h_init[i] = P.Dv(CD_HINIT, i);
  if (h_final) h_final[i] = P.Dv(CD_HLAST, i);
}

// ------------------------------------------------------------------------ host driver
template <class M, int SOLVER>
int phased_leapfrog_steps(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld,
                          const Opts& o, const StepsIO& io, cudaStream_t s) {
  constexpr int U = M::U;
  PhasedLayout L;
  L.lds = pad32(n), L.U = U, L.Y = D.Y, L.NGND = M::NG * M::ND;
  ScratchBuf sb;
  int rc = sb.alloc(L.rows_total(), L.lds, s);
  if (rc != MLB_OK) return rc;
  L.base = sb.p;
  int* counter = reinterpret_cast<int*>(sb.p + (L.rows_total() - 1) * L.lds);
  // pinned word the solver loop polls; owned by the model handle, so host-driven calls on
  // one handle are serialised
  std::lock_guard<std::mutex> lock(m->pin_mutex);
  if (!m->pinned_word && cudaMallocHost((void**)&m->pinned_word, sizeof(int)) != cudaSuccess) {
    g_err = "pinned counter allocation failed";
    return MLB_ERR_CUDA;
  }
  int* host_count = m->pinned_word;
  const dim3 blk(kStreamBlock), g1(sgrid(n)), g_role(sgrid(n), M::N_ROLE), g_grp(sgrid(n), M::NG);
  auto launched = [&]() { return finish_launch(); };
#define MLB_KP(call)                    \
  do {                                  \
    call;                               \
    if ((rc = launched()) != MLB_OK) return rc; \
  } while (0)
  auto read_counter = [&]() -> int {
    cudaMemcpyAsync(host_count, counter, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemsetAsync(counter, 0, sizeof(int), s);
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    return *host_count;
  };
  cudaMemsetAsync(counter, 0, sizeof(int), s);
  MLB_KP((kp_begin<M><<<g1, blk, 0, s>>>(L, n, ld, io.q, io.p)));
  bool first = true;
  for (int step = 0; step < io.n_step || first; ++step) {
    for (int sub = 0; sub < 3; ++sub) {
      const bool fresh = first || sub == 1;
      if (fresh) {  // evaluate at the working position
        MLB_KP((kp_jacob<M, true><<<g_role, blk, 0, s>>>(D, L, n, ld, io.q)));
        MLB_KP((kp_eval_mid<M><<<g1, blk, 0, s>>>(D, L, n, ld, io.q, io.p)));
        MLB_KP((kp_second_order<M><<<g_grp, blk, 0, s>>>(D, L, n, ld, io.q)));
      }
      if (io.n_step <= 0) {  // only h_init is wanted
        MLB_KP((kp_kick_project<M><<<g1, blk, 0, s>>>(L, n, ld, io.p, io.q, io.dt, io.dt_pc,
                                                      1, 0, 1, 0)));
        first = false;
        break;
      }
      // note: with first == true the Hamiltonian is recorded BEFORE the kick
      MLB_KP((kp_kick_project<M><<<g1, blk, 0, s>>>(L, n, ld, io.p, io.q, io.dt, io.dt_pc,
                                                    fresh, sub != 1, first, sub == 2)));
      first = false;
      if (sub == 2) break;
      const double sign = sub == 0 ? 1.0 : -1.0;
      MLB_KP((kp_solve_begin<M><<<g1, blk, 0, s>>>(L, n, ld, io.q, io.p, io.dt, io.dt_pc, sign,
                                                   o.max_iters, counter)));
      int active = read_counter();
      while (active > 0) {
        if (SOLVER == MLB_SOLVER_QUASI_NEWTON)
          MLB_KP((kp_constr<M><<<g1, blk, 0, s>>>(D, L, n)));
        else
          MLB_KP((kp_jacob<M, false><<<g_role, blk, 0, s>>>(D, L, n, ld, io.q)));
        MLB_KP((kp_update<M, SOLVER><<<g1, blk, 0, s>>>(L, n, o, counter)));
        active = read_counter();
      }
      if (active < 0) {
        g_err = "stream synchronisation failed in the solver loop";
        return MLB_ERR_CUDA;
      }
      MLB_KP((kp_solve_end<M><<<g1, blk, 0, s>>>(L, n, ld, io.q, io.p, io.dt, io.dt_pc, sign, o,
                                                 sub == 0)));
    }
    if (io.n_step <= 0) break;
  }
  MLB_KP((kp_finish<M><<<g1, blk, 0, s>>>(L, n, io.status, io.n_done, io.iters_fwd,
                                          io.iters_rev, io.h_init, io.h_final)));
#undef MLB_KP
  return MLB_OK;
}

}  // namespace mlb
