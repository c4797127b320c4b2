// Kernels + launcher table for the streaming (large dim_y, ODE-constrained) family; the
// C ABI entry points of include/mlb.h dispatch here through the same ModelOps table as
// the register-resident models.  One chain per thread, one warp per CTA so that a batch
// of a few thousand chains still spreads over all 148 SMs.
#pragma once
#include <mutex>

#include "mlb_launch.cuh"
#include "mlb_stream.cuh"

namespace mlb {

constexpr int kStreamBlock = 32;

#define MLB_STREAM_INDEX()                                          \
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
  if (i >= n) return;

MLB_DEV Col col(double* p, int64_t ld, int64_t i) { return Col{p ? p + i : nullptr, ld}; }
MLB_DEV Col col(const double* p, int64_t ld, int64_t i) {
  return Col{p ? const_cast<double*>(p) + i : nullptr, ld};
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    ks_constr(const __grid_constant__ OdeData D, int64_t n, int64_t ld, const double* q, double* c) {
  MLB_STREAM_INDEX();
  const Col qc = col(q, ld, i);
  double u[M::U];
  Stream<M>::load_u(qc, u);
  Stream<M>::constr(D, u, [&](int k) { return qc[M::U + k]; }, col(c, ld, i), MLB_NORM_MAX);
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    ks_jacob(const __grid_constant__ OdeData D, int64_t n, int64_t ld, const double* q, double* jac, double* c) {
  MLB_STREAM_INDEX();
  Stream<M>::jacob(D, col(q, ld, i), col(jac, ld, i), col(c, ld, i), MLB_NORM_MAX);
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    ks_decompose(int Y, int64_t n, int64_t ld, const double* jac, double* gram) {
  MLB_STREAM_INDEX();
  typename Stream<M>::Chol C;
  Stream<M>::decompose(Y, col(jac, ld, i), col(gram, ld, i), C);
}

// scratch: [jac_rows + gram_rows + Q + U*Y][lds]
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    ks_logdet(const __grid_constant__ OdeData D, int64_t n, int64_t ld, const double* q, double* val, double* grad,
              double* c, double* jac, double* gram, double* nld_val, double* nld_grad,
              double* h1, double* dh1, double* scratch, int64_t lds) {
  MLB_STREAM_INDEX();
  using S = Stream<M>;
  constexpr int U = M::U;
  const int Y = D.Y, Q = U + Y, jac_rows = U * Y + Y, gram_rows = U * U + Y;
  const Col qc = col(q, ld, i);
  Col sc = col(scratch, lds, i);
  const Col J = jac ? col(jac, ld, i) : sc.rows(0);
  const Col G = gram ? col(gram, ld, i) : sc.rows(jac_rows);
  const Col gtot = sc.rows(jac_rows + gram_rows), B = sc.rows(jac_rows + gram_rows + Q);
  typename S::Chol C;
  const bool want_grad = grad || nld_val || nld_grad || h1 || dh1;
  double ldv, nv = 0.0;
  if (want_grad) {
    const typename S::EvalOut e = S::evaluate(D, qc, J, G, gtot, B, C, col(c, ld, i));
    ldv = e.logdet, nv = e.nld;
    // gtot = grad nld + grad logdet; grad nld = (u - m, n): split them again
    double u[U], gn[U];
    S::load_u(qc, u);
    M::nld_u(u, gn);
#pragma unroll
    for (int a = 0; a < U; ++a) {
      const double gt = gtot[a];
      if (grad) grad[(int64_t)a * ld + i] = gt - gn[a];
      if (nld_grad) nld_grad[(int64_t)a * ld + i] = gn[a];
      if (dh1) dh1[(int64_t)a * ld + i] = gt;
    }
#pragma unroll 1
    for (int k = U; k < Q; ++k) {
      const double gnk = qc[k], gt = gtot[k];
      if (grad) grad[(int64_t)k * ld + i] = gt - gnk;
      if (nld_grad) nld_grad[(int64_t)k * ld + i] = gnk;
      if (dh1) dh1[(int64_t)k * ld + i] = gt;
    }
  } else {
    S::jacob(D, qc, J, col(c, ld, i), MLB_NORM_MAX);
    S::decompose(Y, J, G, C);
    ldv = S::logdet(Y, G, C);
  }
  if (val) val[i] = ldv;
  if (nld_val) nld_val[i] = nv;
  if (h1) h1[i] = nv + ldv;
}

template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    ks_nld(const __grid_constant__ OdeData D, int64_t n, int64_t ld, const double* q, double* val, double* grad) {
  MLB_STREAM_INDEX();
  constexpr int U = M::U;
  const Col qc = col(q, ld, i);
  double u[U], g[U];
  Stream<M>::load_u(qc, u);
  double v = M::nld_u(u, g);
#pragma unroll
  for (int a = 0; a < U; ++a)
    if (grad) grad[(int64_t)a * ld + i] = g[a];
#pragma unroll 1
  for (int k = 0; k < D.Y; ++k) {
    const double nk = qc[U + k];
    v = fma(0.5 * nk, nk, v);
    if (grad) grad[(int64_t)(U + k) * ld + i] = nk;
  }
  if (val) val[i] = v;
}

// scratch: [Y][lds] (used by OP_NSC / OP_PINV for the intermediate Y-vector)
template <class M, int OP>
__global__ void __launch_bounds__(kStreamBlock)
    ks_block_op(int Y, int64_t n, int64_t ld, const double* jac, const double* jac_r,
                const double* gram, const double* vec, double* out, double* scratch,
                int64_t lds) {
  MLB_STREAM_INDEX();
  using S = Stream<M>;
  constexpr int U = M::U;
  const Col J = col(jac, ld, i), v = col(vec, ld, i), o = col(out, ld, i);
  const Col w = col(scratch, lds, i);
  if (OP == OP_LMULT) {
    S::lmult(Y, J, v, o);
  } else if (OP == OP_RMULT) {
    S::rmult(Y, J, v, o);
  } else if (OP == OP_PINV || OP == OP_NSC) {
    typename S::Chol C;
    S::load_chol(col(gram, ld, i), C);
    if (OP == OP_NSC) {
      S::lmult(Y, J, v, w);
    } else {
#pragma unroll 1
      for (int k = 0; k < Y; ++k) w[k] = v[k];
    }
    S::template inv_gram<true, false>(Y, J, C, w, o);
  } else {
    const Col Jr = col(jac_r, ld, i);
    double s[U];
    const bool ok = S::inv_jacprod_core(Y, J, Jr, v, s);
#pragma unroll 1
    for (int k = 0; k < Y; ++k)
      o[k] = ok ? S::inv_jacprod_x(Y, k, J, Jr, v[k], s) : nan("");
  }
}

// scratch: [jac_rows + gram_rows + Y][lds]
template <class M>
__global__ void __launch_bounds__(kStreamBlock)
    ks_project_cotangent(const __grid_constant__ OdeData D, int64_t n, int64_t ld, const double* q, double* p,
                         double* scratch, int64_t lds) {
  MLB_STREAM_INDEX();
  using S = Stream<M>;
  const int Y = D.Y, jac_rows = M::U * Y + Y, gram_rows = M::U * M::U + Y;
  Col sc = col(scratch, lds, i);
  const Col J = sc.rows(0), G = sc.rows(jac_rows), w = sc.rows(jac_rows + gram_rows);
  typename S::Chol C;
  S::jacob(D, col(q, ld, i), J, Col{nullptr, 0}, MLB_NORM_MAX);
  S::decompose(Y, J, G, C);
  S::project(Y, J, C, col(p, ld, i), w);
}

// scratch: [jac_rows + 2 Y + Q][lds]  (Jt, c, dir, mu)
template <class M, int SOLVER>
__global__ void __launch_bounds__(kStreamBlock)
    ks_solve(const __grid_constant__ OdeData D, int64_t n, int64_t ld, const double* q, const double* jac_prev,
             const double* gram_prev, double dt, Opts o, double* q_out, double* mu_out,
             int32_t* iters, double* norm_dq, double* err, int32_t* n_constr,
             int32_t* status, double* scratch, int64_t lds) {
  MLB_STREAM_INDEX();
  using S = Stream<M>;
  constexpr int U = M::U;
  const int Y = D.Y, Q = U + Y, jac_rows = U * Y + Y;
  Col sc = col(scratch, lds, i);
  const Col jt = sc.rows(0), c = sc.rows(jac_rows), dir = sc.rows(jac_rows + Y);
  const Col mu = mu_out ? col(mu_out, ld, i) : sc.rows(jac_rows + 2 * Y);
  const Col qo = col(q_out, ld, i), qi = col(q, ld, i), jp = col(jac_prev, ld, i);
  if (q_out != q) {
#pragma unroll 1
    for (int k = 0; k < Q; ++k) qo[k] = qi[k];
  }
  typename S::Chol Cp;
  Cp.ok = true;
  if (SOLVER == MLB_SOLVER_QUASI_NEWTON) {
    if (gram_prev) S::load_chol(col(gram_prev, ld, i), Cp);
    else S::decompose(Y, jp, Col{nullptr, 0}, Cp);
  }
  const SolveOut r = S::template solve<SOLVER>(D, qo, jp, Cp, dt, o, mu, jt, c, dir);
  if (iters) iters[i] = r.iters;
  if (norm_dq) norm_dq[i] = r.ndq;
  if (err) err[i] = r.err;
  if (n_constr) n_constr[i] = r.n_constr;
  status[i] = r.status;
}

// scratch: [StreamScratch::rows][lds]
template <class M, int SOLVER>
__global__ void __launch_bounds__(kStreamBlock)
    ks_leapfrog_steps(const __grid_constant__ OdeData D, int64_t n, int64_t ld, double* q, double* p, double dt,
                      const double* dt_pc, int n_step, Opts o, int32_t* status,
                      int32_t* n_done, int32_t* iters_fwd, int32_t* iters_rev, double* h_init,
                      double* h_final, double* scratch, int64_t lds) {
  MLB_STREAM_INDEX();
  const StreamScratch S = StreamScratch::make(scratch, lds, i, M::U, D.Y);
  const double dti = dt_pc ? dt_pc[i] : dt;
  const TrajOut t = Stream<M>::template trajectory<SOLVER, false, false>(
      D, col(q, ld, i), col(p, ld, i), S, n_step, dti, o, 0.0);
  status[i] = t.status;
  if (n_done) n_done[i] = t.done;
  if (iters_fwd) iters_fwd[i] = t.it_fwd;
  if (iters_rev) iters_rev[i] = t.it_rev;
  if (h_init) h_init[i] = t.h_init;
  if (h_final) h_final[i] = t.h_last;
}

// scratch: [StreamScratch::rows + 2 Q][lds]  (proposal position / momentum)
template <class M, int SOLVER>
__global__ void __launch_bounds__(kStreamBlock)
    ks_hmc_sample(const __grid_constant__ OdeData D, int64_t n, int64_t ld, double* q, Opts o,
                  SampleArgs a, double* scratch, int64_t lds) {
  MLB_STREAM_INDEX();
  constexpr int U = M::U;
  const int Q = U + D.Y;
  const StreamScratch S = StreamScratch::make(scratch, lds, i, U, D.Y);
  Col sc = col(scratch, lds, i);
  const Col qc = sc.rows((int)StreamScratch::rows(U, D.Y)), pc = qc.rows(Q);
  const Col qv = col(q, ld, i);
  double ad_iter = 0, ad_smooth = 0, ad_err = 0, ad_target = 0, dti = a.dt;
  if (a.adapt) {
    ad_iter = a.adapt[0 * ld + i], ad_smooth = a.adapt[1 * ld + i];
    ad_err = a.adapt[2 * ld + i], ad_target = a.adapt[3 * ld + i];
    dti = a.adapt[4 * ld + i];
  }
  double s_acc = 0, s_nacc = 0, s_conv = 0, s_nonrev = 0, s_hdiv = 0, s_steps = 0, s_its = 0;
  int last = MLB_ST_OK;
#pragma unroll 1
  for (int it = 0; it < a.n_iter; ++it) {
    if (a.mom_noise) {
      const double* mn = a.mom_noise + (int64_t)it * Q * ld + i;
#pragma unroll 1
      for (int k = 0; k < Q; ++k) pc[k] = mn[(int64_t)k * ld];
    } else {
#pragma unroll 1
      for (int j = 0; 2 * j < Q; ++j) {
        double n0, n1;
        philox_normal_pair(a.seed, a.chain0 + (uint64_t)i, a.iter0 + (uint32_t)it,
                           (uint32_t)j, n0, n1);
        pc[2 * j] = n0;
        if (2 * j + 1 < Q) pc[2 * j + 1] = n1;
      }
    }
#pragma unroll 1
    for (int k = 0; k < Q; ++k) qc[k] = qv[k];
    const double u = a.unif ? a.unif[(int64_t)it * ld + i]
                            : philox_uniform(a.seed, a.chain0 + (uint64_t)i,
                                             a.iter0 + (uint32_t)it);
    const TrajOut t = Stream<M>::template trajectory<SOLVER, true, true>(
        D, qc, pc, S, a.n_step, dti, o, a.max_delta_h);
    const int st = t.status;
    s_its += t.it_fwd + t.it_rev;
    s_steps += t.done;
    double a_stat = 0.0;
    if (st == MLB_ST_OK) {
      const double d = t.h_init - t.h_last;
      a_stat = (d != d) ? 0.0 : exp(d < 0.0 ? d : 0.0);
      if (u < a_stat) {
#pragma unroll 1
        for (int k = 0; k < Q; ++k) qv[k] = qc[k];
        s_nacc += 1;
      }
    }
    s_acc += a_stat;
    s_conv += (st == MLB_ST_DIVERGED || st == MLB_ST_NOT_CONVERGED || st == MLB_ST_LINALG);
    s_nonrev += (st == MLB_ST_NON_REVERSIBLE);
    s_hdiv += (st == MLB_ST_H_DIVERGED);
    last = st;
    if (a.trace_accept) a.trace_accept[(int64_t)it * ld + i] = a_stat;
    if (a.trace_q && (it + 1) % a.trace_every == 0) {
      double* tq = a.trace_q + (int64_t)((it + 1) / a.trace_every - 1) * a.trace_dim * ld + i;
#pragma unroll 1
      for (int k = 0; k < a.trace_dim; ++k) tq[(int64_t)k * ld] = qv[k];
    }
    if (a.adapt) {  // Mici DualAveragingStepSizeAdapter.update
      ad_iter += 1;
      const double w = 1.0 / (a.a_offset + ad_iter);
      ad_err = (1.0 - w) * ad_err + w * (a.a_target - a_stat);
      const double sw = pow(1.0 / ad_iter, a.a_decay);
      const double log_eps = ad_target - ad_err * sqrt(ad_iter) / a.a_reg;
      ad_smooth = (1.0 - sw) * ad_smooth + sw * log_eps;
      dti = exp(log_eps);
    }
  }
  if (a.adapt) {
    a.adapt[0 * ld + i] = ad_iter, a.adapt[1 * ld + i] = ad_smooth;
    a.adapt[2 * ld + i] = ad_err, a.adapt[4 * ld + i] = dti;
  }
  if (a.stats) {
    a.stats[0 * ld + i] += s_acc, a.stats[1 * ld + i] += s_nacc;
    a.stats[2 * ld + i] += s_conv, a.stats[3 * ld + i] += s_nonrev;
    a.stats[4 * ld + i] += s_hdiv, a.stats[5 * ld + i] += s_steps;
    a.stats[6 * ld + i] += s_its;
  }
  if (a.last_status) a.last_status[i] = last;
}

// ----------------------------------------------------------------------- launchers

// device copy of the model constants, uploaded on first use on the current device
int ode_data(const mlb_model* m, OdeData* out);  // defined in mlb_api.cu

struct ScratchBuf {  // stream-ordered scratch allocation
  double* p = nullptr;
  cudaStream_t s;
  int alloc(int64_t rows, int64_t lds, cudaStream_t stream) {
    s = stream;
    // (library-owned pool whose freed blocks stay cached: mlb_launch.cuh)
    cudaError_t e = scratch_alloc((void**)&p, sizeof(double) * (size_t)rows * (size_t)lds, s);
    if (e != cudaSuccess) {
      g_err = std::string("scratch allocation: ") + cudaGetErrorString(e);
      p = nullptr;
      return MLB_ERR_CUDA;
    }
    return MLB_OK;
  }
  ~ScratchBuf() {
    if (p) cudaFreeAsync(p, s);
  }
};

inline unsigned sgrid(int64_t n) { return (unsigned)((n + kStreamBlock - 1) / kStreamBlock); }
inline int64_t pad32(int64_t n) { return (n + 31) / 32 * 32; }

// host-driven sub-phase path (mlb_stream_phased.cuh)
template <class M, int SOLVER>
int phased_leapfrog_steps(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld,
                          const Opts& o, const StepsIO& io, cudaStream_t s);

template <class M, int SOLVER>
int phased_hmc_sample(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld, double* q,
                      const Opts& o, const SampleArgs& a, cudaStream_t s);

template <class M, int SOLVER>
int phased_nuts_sample(const mlb_model* m, const OdeData& D, int64_t n, int64_t ld, double* q,
                       const Opts& o, const SampleArgs& a, const NutsArgs& na, cudaStream_t s);

template <class M>
struct StreamLaunch {
  static constexpr int U = M::U;
#define MLB_ODE()                  \
  OdeData D;                       \
  {                                \
    int rc_ = ode_data(m, &D);     \
    if (rc_ != MLB_OK) return rc_; \
  }                                \
  const int Y = D.Y;               \
  (void)Y
  static int constr(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* c,
                    cudaStream_t s) {
    MLB_ODE();
    ks_constr<M><<<sgrid(n), kStreamBlock, 0, s>>>(D, n, ld, q, c);
    return finish_launch();
  }
  static int jacob(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* jac,
                   double* c, cudaStream_t s) {
    MLB_ODE();
    ks_jacob<M><<<sgrid(n), kStreamBlock, 0, s>>>(D, n, ld, q, jac, c);
    return finish_launch();
  }
  static int decompose(const mlb_model* m, int64_t n, int64_t ld, const double* jac,
                       double* gram, cudaStream_t s) {
    ks_decompose<M><<<sgrid(n), kStreamBlock, 0, s>>>(m->Y, n, ld, jac, gram);
    return finish_launch();
  }
  static int logdet(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* val,
                    double* grad, double* c, double* jac, double* gram, double* nld_val,
                    double* nld_grad, double* h1, double* dh1, cudaStream_t s) {
    MLB_ODE();
    ScratchBuf sb;
    const int64_t lds = pad32(n);
    int rc = sb.alloc((int64_t)m->jac_rows + m->gram_rows + m->Q + U * Y, lds, s);
    if (rc != MLB_OK) return rc;
    ks_logdet<M><<<sgrid(n), kStreamBlock, 0, s>>>(D, n, ld, q, val, grad, c, jac, gram,
                                                  nld_val, nld_grad, h1, dh1, sb.p, lds);
    return finish_launch();
  }
  static int nld(const mlb_model* m, int64_t n, int64_t ld, const double* q, double* val,
                 double* grad, cudaStream_t s) {
    MLB_ODE();
    ks_nld<M><<<sgrid(n), kStreamBlock, 0, s>>>(D, n, ld, q, val, grad);
    return finish_launch();
  }
  static int block_op(const mlb_model* m, int op, int64_t n, int64_t ld, const double* jac,
                      const double* jac_r, const double* gram, const double* vec,
                      double* out, cudaStream_t s) {
    const int Y = m->Y;
    ScratchBuf sb;
    const int64_t lds = pad32(n);
    int rc = sb.alloc(Y, lds, s);
    if (rc != MLB_OK) return rc;
    const unsigned g = sgrid(n);
    switch (op) {
      case OP_LMULT: ks_block_op<M, OP_LMULT><<<g, kStreamBlock, 0, s>>>(Y, n, ld, jac, jac_r, gram, vec, out, sb.p, lds); break;
      case OP_RMULT: ks_block_op<M, OP_RMULT><<<g, kStreamBlock, 0, s>>>(Y, n, ld, jac, jac_r, gram, vec, out, sb.p, lds); break;
      case OP_PINV: ks_block_op<M, OP_PINV><<<g, kStreamBlock, 0, s>>>(Y, n, ld, jac, jac_r, gram, vec, out, sb.p, lds); break;
      case OP_NSC: ks_block_op<M, OP_NSC><<<g, kStreamBlock, 0, s>>>(Y, n, ld, jac, jac_r, gram, vec, out, sb.p, lds); break;
      default: ks_block_op<M, OP_INVJP><<<g, kStreamBlock, 0, s>>>(Y, n, ld, jac, jac_r, gram, vec, out, sb.p, lds); break;
    }
    return finish_launch();
  }
  static int project_cotangent(const mlb_model* m, int64_t n, int64_t ld, const double* q,
                               double* p, cudaStream_t s) {
    MLB_ODE();
    ScratchBuf sb;
    const int64_t lds = pad32(n);
    int rc = sb.alloc((int64_t)m->jac_rows + m->gram_rows + Y, lds, s);
    if (rc != MLB_OK) return rc;
    ks_project_cotangent<M><<<sgrid(n), kStreamBlock, 0, s>>>(D, n, ld, q, p, sb.p, lds);
    return finish_launch();
  }
  static int solve(const mlb_model* m, int solver, int64_t n, int64_t ld, const Opts& o,
                   const SolveIO& io, cudaStream_t s) {
    MLB_ODE();
    if (solver > MLB_SOLVER_NEWTON_LS) {
      g_err = "Mici's own solvers (MLB_SOLVER_MICI_*) are only built for the toy / "
              "hierarchical models; ODE models use the mlift solvers as the reference does";
      return MLB_ERR_UNSUPPORTED;
    }
    ScratchBuf sb;
    const int64_t lds = pad32(n);
    int rc = sb.alloc((int64_t)m->jac_rows + 2 * Y + m->Q, lds, s);
    if (rc != MLB_OK) return rc;
    auto go = [&]<int S>() {
      ks_solve<M, S><<<sgrid(n), kStreamBlock, 0, s>>>(
          D, n, ld, io.q, io.jac_prev, io.gram_prev, io.dt, o, io.q_out, io.mu_out, io.iters,
          io.norm_dq, io.err, io.n_constr, io.status, sb.p, lds);
      return finish_launch();
    };
    if (solver == MLB_SOLVER_QUASI_NEWTON) return go.template operator()<MLB_SOLVER_QUASI_NEWTON>();
    if (solver == MLB_SOLVER_NEWTON) return go.template operator()<MLB_SOLVER_NEWTON>();
    return go.template operator()<MLB_SOLVER_NEWTON_LS>();
  }
  static int steps(const mlb_model* m, int solver, int64_t n, int64_t ld, const Opts& o,
                   const StepsIO& io, cudaStream_t s) {
    MLB_ODE();
    if (solver > MLB_SOLVER_NEWTON_LS) {
      g_err = "Mici's own solvers are not built for ODE models";
      return MLB_ERR_UNSUPPORTED;
    }
    if (!(o.flags & MLB_FLAG_SINGLE_KERNEL) && solver != MLB_SOLVER_NEWTON_LS) {
      if (solver == MLB_SOLVER_QUASI_NEWTON)
        return phased_leapfrog_steps<M, MLB_SOLVER_QUASI_NEWTON>(m, D, n, ld, o, io, s);
      return phased_leapfrog_steps<M, MLB_SOLVER_NEWTON>(m, D, n, ld, o, io, s);
    }
    ScratchBuf sb;
    const int64_t lds = pad32(n);
    int rc = sb.alloc(StreamScratch::rows(U, Y), lds, s);
    if (rc != MLB_OK) return rc;
    auto go = [&]<int S>() {
      ks_leapfrog_steps<M, S><<<sgrid(n), kStreamBlock, 0, s>>>(
          D, n, ld, io.q, io.p, io.dt, io.dt_pc, io.n_step, o, io.status, io.n_done,
          io.iters_fwd, io.iters_rev, io.h_init, io.h_final, sb.p, lds);
      return finish_launch();
    };
    if (solver == MLB_SOLVER_QUASI_NEWTON) return go.template operator()<MLB_SOLVER_QUASI_NEWTON>();
    if (solver == MLB_SOLVER_NEWTON) return go.template operator()<MLB_SOLVER_NEWTON>();
    return go.template operator()<MLB_SOLVER_NEWTON_LS>();
  }
  static int sample(const mlb_model* m, int solver, int64_t n, int64_t ld, double* q,
                    const Opts& o, const SampleArgs& a, cudaStream_t s) {
    MLB_ODE();
    if (solver > MLB_SOLVER_NEWTON_LS) {
      g_err = "Mici's own solvers are not built for ODE models";
      return MLB_ERR_UNSUPPORTED;
    }
    if (!(o.flags & MLB_FLAG_SINGLE_KERNEL) && solver != MLB_SOLVER_NEWTON_LS) {
      if (solver == MLB_SOLVER_QUASI_NEWTON)
        return phased_hmc_sample<M, MLB_SOLVER_QUASI_NEWTON>(m, D, n, ld, q, o, a, s);
      return phased_hmc_sample<M, MLB_SOLVER_NEWTON>(m, D, n, ld, q, o, a, s);
    }
    ScratchBuf sb;
    const int64_t lds = pad32(n);
    int rc = sb.alloc(StreamScratch::rows(U, Y) + 2 * (int64_t)m->Q, lds, s);
    if (rc != MLB_OK) return rc;
    auto go = [&]<int S>() {
      ks_hmc_sample<M, S><<<sgrid(n), kStreamBlock, 0, s>>>(D, n, ld, q, o, a, sb.p, lds);
      return finish_launch();
    };
    if (solver == MLB_SOLVER_QUASI_NEWTON) return go.template operator()<MLB_SOLVER_QUASI_NEWTON>();
    if (solver == MLB_SOLVER_NEWTON) return go.template operator()<MLB_SOLVER_NEWTON>();
    return go.template operator()<MLB_SOLVER_NEWTON_LS>();
  }
  static int nuts(const mlb_model* m, int solver, int64_t n, int64_t ld, double* q,
                  const Opts& o, const SampleArgs& a, const NutsArgs& na, cudaStream_t s) {
    MLB_ODE();
    if (solver > MLB_SOLVER_NEWTON) {
      g_err = "the host-driven dynamic sampler takes the quasi-Newton and Newton solvers";
      return MLB_ERR_UNSUPPORTED;
    }
    if (solver == MLB_SOLVER_QUASI_NEWTON)
      return phased_nuts_sample<M, MLB_SOLVER_QUASI_NEWTON>(m, D, n, ld, q, o, a, na, s);
    return phased_nuts_sample<M, MLB_SOLVER_NEWTON>(m, D, n, ld, q, o, a, na, s);
  }
#undef MLB_ODE
};

template <class M>
ModelOps make_stream_ops(int model, int param) {
  using L = StreamLaunch<M>;
  return ModelOps{model, 0 /* any dim_y */, param, &L::constr, &L::jacob, &L::decompose,
                  &L::logdet, &L::nld, &L::block_op, &L::project_cotangent, &L::solve,
                  &L::steps, &L::sample, &L::nuts};
}

}  // namespace mlb
